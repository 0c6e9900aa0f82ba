"""Proofs/s of the device-resident TurboPlonK prover on synthetic circuits (BASELINE configs[4]), with the per-round split.

    python scripts/gpu_plonk_bench.py [log_sizes ...]      default: 13 14 18 20 22
"""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi, plonk, KZGCommitmentSchemeBN254
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript

ffi.init(0)
sizes = [int(a) for a in sys.argv[1:]] or [13, 14, 18, 20, 22]
TAU = plonk.mont(0x1234567890ABCDEF1234567890ABCDEF)
rows = []
for lg in sizes:
    n = 1 << lg
    t0 = time.perf_counter()
    cs = plonk.TurboCS.synthetic(lg)
    t_cs = time.perf_counter() - t0
    t0 = time.perf_counter()
    pcs = KZGCommitmentSchemeBN254.new(n + 2, TAU)
    t_srs = time.perf_counter() - t0
    t0 = time.perf_counter()
    params = plonk.indexer(cs, pcs)
    torch.cuda.synchronize()
    t_idx = time.perf_counter() - t0
    wit = plonk.DevVec.from_numpy(cs.get_witness_array(), torch.device("cuda", 0))
    reps = 5 if lg <= 18 else 3
    for _ in range(2):
        plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit)
    torch.cuda.synchronize()
    timings = {}
    launches0 = ffi.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit, timings=timings)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    row = {"log_n": lg, "prove_ms": round(dt * 1e3, 2), "proofs_per_s": round(1 / dt, 3), "launches_per_proof": (ffi.launch_count() - launches0) // reps,
           "rounds_ms": {k: round(v / reps, 2) for k, v in timings.items()}, "setup_s": {"circuit": round(t_cs, 2), "srs": round(t_srs, 2), "indexer": round(t_idx, 2)},
           "srs": {k: pcs.info()[k] for k in ("window_bits", "windows", "batch_slots")}, "hbm_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    pcs.close()
    del params, wit, cs, pcs
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/plonk_bench.json", "w"), indent=1)
