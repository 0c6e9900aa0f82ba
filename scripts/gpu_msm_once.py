"""One MSM configuration, a few launches (for ncu): python scripts/gpu_msm_once.py LOG_N [C]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
import bench as B
lg = int(sys.argv[1]); c = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ffi.init(0)
if os.environ.get('L2_FETCH'):
    ffi.configure('l2_fetch_granularity', int(os.environ['L2_FETCH']))
dev = torch.device("cuda", 0)
n = 1 << lg
bases = ffi.srs_generate(B.random_fr(1, 5)[0], n)
h = ffi.srs_upload(bases, c)
sc = torch.from_numpy(B.random_fr(n, 2).view(np.int64)).to(dev)
out = torch.zeros(12, dtype=torch.int64, device=dev)
for _ in range(4):
    ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr())
e1.record()
torch.cuda.synchronize()
print("ms per MSM", e0.elapsed_time(e1) / 10, "L2_FETCH", os.environ.get("L2_FETCH"))
print(ffi.srs_info(h))
