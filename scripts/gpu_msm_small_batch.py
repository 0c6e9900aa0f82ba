"""The prover's MSM shape in isolation for an ncu capture: a batch of 5 MSMs of 2^14 points, once with 2 lanes per bucket (the old
rule's choice) and once with the engine's rule (8).

    ncu --set full --clock-control none -k regex:msm_accumulate --launch-skip 4 -c 2 -o out python scripts/gpu_msm_small_batch.py
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
n, k = 1 << 14, 5
h = ffi.srs_upload(ffi.srs_generate(B.random_fr(1, 5)[0], n), 0)
sc = [torch.from_numpy(B.random_fr(n, 10 + j).view(np.int64)).to(dev) for j in range(k)]
out = torch.zeros(12 * k, dtype=torch.int64, device=dev)
for lanes in (2, 0, 2, 0, 2, 0):          # two warm-up pairs (4 accumulate launches to skip), then the captured pair
    ffi.configure("msm_lanes", lanes)
    ffi.msm_g1_batch_device(h, [s.data_ptr() for s in sc], [n] * k, out.data_ptr())
    torch.cuda.synchronize()
