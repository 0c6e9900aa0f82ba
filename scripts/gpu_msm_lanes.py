"""Lanes per bucket of the MSM accumulate kernel (uzkge_cuda_configure("msm_lanes", g); 0 = the engine's rule) against MSM time:
single MSMs 2^12..2^20 and the prover's batch shapes (8 and 5 MSMs of 2^14).  Device-resident, CUDA events.

    python scripts/gpu_msm_lanes.py [max_log]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench as B
from uzkge_b200 import ffi

ffi.init(0)
dev = torch.device("cuda", 0)
max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 20
tau = B.random_fr(1, 5)[0]


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for lg in range(12, max_log + 1, 2):
    n = 1 << lg
    h = ffi.srs_upload(ffi.srs_generate(tau, n), 0)
    info = ffi.srs_info(h)
    sc = [torch.from_numpy(B.random_fr(n, 10 + j).view(np.int64)).to(dev) for j in range(8)]
    out = torch.zeros(12 * 8, dtype=torch.int64, device=dev)
    shapes = [1] + ([5, 8] if lg <= 16 else [])
    for k in shapes:
        row = {}
        for lanes in (0, 1, 2, 4, 8, 16, 32):
            ffi.configure("msm_lanes", lanes)
            if k == 1:
                row[lanes] = timeit(lambda: ffi.msm_g1_device(h, sc[0].data_ptr(), n, out.data_ptr()))
            else:
                row[lanes] = timeit(lambda: ffi.msm_g1_batch_device(h, [s.data_ptr() for s in sc[:k]], [n] * k, out.data_ptr()))
        ffi.configure("msm_lanes", 0)
        print(f"2^{lg} c={info['window_bits']} batch={k}: " + "  ".join(f"g{g}={t:.0f}us" for g, t in row.items()), flush=True)
    ffi.srs_free(h)
