"""The kernels the device group adds to the quotient round, alone on one GPU at n = 2^LOG (for ncu): one coset of the quotient map
(uzkge_cuda_plonk_quotient_range_fr_device), the strided copies between a coset and a compact vector, the size-n coset inverse
transform and the 6-point combine over the cosets (uzkge_cuda_plonk_coset_combine_fr_device).  Checks the round trip
coefficients -> coset values -> per-coset inverse transforms -> combine == coefficients before exiting 0.
    python scripts/gpu_coset_kernels_once.py [LOG]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench as B
from uzkge_b200 import ffi, plonk

ffi.init(0)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, factor = 1 << lg, 6
m = n * factor
dev = torch.device("cuda", 0)
FR = plonk.FR_MODULUS
k1 = 0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A
k1_m = plonk.mont_rows([k1])[0]
t = torch.from_numpy(B.random_fr(m, 9).view(np.int64)).to(dev)
ev, scr = torch.empty_like(t), torch.empty_like(t)
ffi.ntt_fr_device(t.data_ptr(), ev.data_ptr(), scr.data_ptr(), m, m, False, k1_m)          # t on the coset k1 <w_m>
w_m = int.from_bytes(bytes(ffi.fr_root_of_unity(m).tobytes()), "little") * pow(1 << 256, -1, FR) % FR
u = torch.empty_like(t)
tmp, s2 = torch.empty(4 * n, dtype=torch.int64, device=dev), torch.empty(4 * n, dtype=torch.int64, device=dev)
for rep in range(2):
    for j in range(factor):
        g_inv = plonk.mont_rows([pow(k1 * pow(w_m, j, FR) % FR, -1, FR)])[0]
        ffi.fr_strided_copy_device(ev.data_ptr(), j, factor, tmp.data_ptr(), 0, 1, n)
        ffi.ntt_fr_device(tmp.data_ptr(), u.data_ptr() + 32 * j * n, s2.data_ptr(), n, n, True, g_inv)
    out = torch.empty_like(t)
    ffi.plonk_coset_combine_fr_device(u.data_ptr(), n, factor, k1_m, out.data_ptr())
torch.cuda.synchronize()
ok = bool(torch.equal(out, t))
# one coset of the quotient map over random columns (the whole map for comparison)
cols = [torch.from_numpy(B.random_fr(m, 20 + i).view(np.int64)).to(dev) for i in range(28)]
p = [c.data_ptr() for c in cols]
sc = B.random_fr(6, 99)
zh = B.random_fr(factor, 98)
q_out = torch.empty_like(t)
kk = plonk.mont_rows([1, k1, 3, 5, 7])
for rng_ in (None, (1, factor, n)):
    for _ in range(2):
        ffi.plonk_quotient_fr_device(p[0:5], p[5:14], p[14], p[15], p[16:21], p[21], p[22], p[23], p[24:28], kk, sc[0], sc[1], sc[2], sc[3],
                                     sc[4], zh, m, factor, q_out.data_ptr(), point_range=rng_)
torch.cuda.synchronize()
print("coset round trip at n = 2^%d: %s" % (lg, "ok" if ok else "MISMATCH"))
sys.exit(0 if ok else 1)
