"""cProfile of the host side of one small proof (where do the 20 ms at n = 2^14 go?)."""
import sys, os, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from uzkge_b200 import ffi, plonk, KZGCommitmentSchemeBN254
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript

ffi.init(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 14
cs = plonk.TurboCS.synthetic(lg)
pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(12345))
params = plonk.indexer(cs, pcs)
wit = plonk.DevVec.from_numpy(cs.get_witness_array(), torch.device("cuda", 0))
for _ in range(3):
    plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
