"""One synthetic proof of 2^LOG gates by the compiled prover, a few times (for ncu): python scripts/gpu_plonk_once.py LOG [REPS]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from uzkge_b200 import KZGCommitmentSchemeBN254, ffi, plonk
from uzkge_b200.native import NativeProver
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript

ffi.init(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cs = plonk.TurboCS.synthetic(lg)
tau = plonk.mont(0x1234567890ABCDEF)
pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, tau)
params = plonk.indexer(cs, pcs)
wit = plonk.DevVec.from_numpy(cs.get_witness_array(), torch.device("cuda", 0))
native = NativeProver(cs, params, pcs)
for _ in range(reps):
    proof = native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), wit)
torch.cuda.synchronize()
print("ok", len(proof.to_bytes_be()), native.last_stats)
