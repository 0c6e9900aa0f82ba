"""One NTT configuration, a few launches (for ncu): python scripts/gpu_ntt_once.py LOG_N"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
import bench as B
lg = int(sys.argv[1])
ffi.init(0)
dev = torch.device("cuda", 0)
n = 1 << lg
x = torch.from_numpy(B.random_fr(n, 2).view(np.int64)).to(dev)
o = torch.empty_like(x); s = torch.empty_like(x)
for _ in range(3):
    ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n)
torch.cuda.synchronize()
print("ok")
