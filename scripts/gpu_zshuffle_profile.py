"""One zshuffle-52 (or, with APP=zmatchmaking, zmatchmaking) proof (production route: everything over the Lagrange SRS) for a kernel
launch list and a host profile.  WINDOW_BITS / MSM_LANES override the engine's choices.

    python scripts/gpu_zshuffle_profile.py host      # cProfile of 10 proofs (where does the host spend the proof's wall time?)
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python scripts/gpu_zshuffle_profile.py ncu   # per-kernel durations of ONE proof (cudaProfilerStart/Stop around it)
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from uzkge_b200 import KZGCommitmentSchemeBN254, ffi, plonk
from uzkge_b200 import shuffle as sh
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript

mode = sys.argv[1] if len(sys.argv) > 1 else "host"
ffi.init(0)
if os.environ.get("MSM_LANES"):
    ffi.configure("msm_lanes", int(os.environ["MSM_LANES"]))
dev = torch.device("cuda", 0)
tau = plonk.mont(0x1234567890ABCDEF1234567890ABCDEF)
prng = ChaChaRng.from_seed(bytes(32))
app = os.environ.get("APP", "zshuffle")          # or "zmatchmaking"
if app == "zshuffle":
    apk = sh.rand_point(prng)
    cs, _ = sh.build_cs(plonk.TurboCS(), prng, apk, [sh.Ciphertext.rand(prng) for _ in range(52)])
    label, count = b"Plonk shuffle Proof", 52
else:
    from uzkge_b200 import matchmaking as mm

    cs, _ = mm.build_cs(plonk.TurboCS(), list(range(1, mm.N + 1)), 12345, 67890)
    apk, label, count = None, mm.PLONK_PROOF_TRANSCRIPT, mm.N
n = cs.size
wb = int(os.environ.get("WINDOW_BITS", "0"))
pcs, lagrange = KZGCommitmentSchemeBN254.new(n + 2, tau, wb), KZGCommitmentSchemeBN254.new_lagrange(n, tau, wb)
params = plonk.indexer(cs, pcs, shuffle=app == "zshuffle", lagrange_pcs=lagrange)
if apk is not None:
    plonk.refresh_prover_params_public_key(cs, params, pcs, apk, lagrange_pcs=lagrange)
wit = plonk.DevVec.from_numpy(cs.get_witness_array(), dev)


native = None
if os.environ.get("PROVER", "native") == "native":       # the compiled prover (uzkge_cuda_plonk_prove); PROVER=mirror: the Python mirror
    from uzkge_b200.native import NativeProver

    native = NativeProver(cs, params, pcs, lagrange, True)


def prove():
    tr = Transcript(label)
    tr.append_u64(count)
    if native is not None:
        return native.prove(ChaChaRng.from_seed(bytes(32)), tr, wit)
    return plonk.prover(ChaChaRng.from_seed(bytes(32)), tr, pcs, cs, params, wit, lagrange_pcs=lagrange, lagrange_all=True)


for _ in range(3):
    prove()
torch.cuda.synchronize()
if mode == "ncu":
    torch.cuda.cudart().cudaProfilerStart()
    prove()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
else:
    t0 = time.perf_counter()
    for _ in range(10):
        prove()
    torch.cuda.synchronize()
    print("ms per proof", (time.perf_counter() - t0) * 100)
    if mode == "time":
        sys.exit(0)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        prove()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
