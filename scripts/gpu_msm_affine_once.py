import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
import bench as B
ffi.init(0)
dev = torch.device("cuda", 0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
bases = ffi.srs_generate(B.random_fr(1, 5)[0], n)
h = ffi.srs_upload(bases, 0)
sc = torch.from_numpy(B.random_fr(n, 2).view(np.int64)).to(dev)
out = torch.zeros(12, dtype=torch.int64, device=dev)
ffi.configure("msm_affine", 1)
for _ in range(2):
    ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr())
torch.cuda.synchronize()
