"""Device-resident timings of the 'next row' kernels: quotient map, Horner scan, grand product, batched MSM."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
import bench as B

ffi.init(0)
dev = torch.device("cuda", 0)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def d(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).to(dev)
peak = ffi.bench_field_mul("fr", 2000)
print(f"Fr mul peak {peak/1e9:.1f} G/s", flush=True)

for logn in (14, 20):
    n, factor = 1 << logn, 6
    m = n * factor
    base = B.random_fr(m, 1)
    arrs = [d(np.roll(base, 7 * i, axis=0)) for i in range(28)]
    sc = B.random_fr(8, 2)
    k = B.random_fr(5, 3)
    zh = B.random_fr(factor, 4)
    out = torch.empty(4 * m, dtype=torch.int64, device=dev)
    p = [t.data_ptr() for t in arrs]
    fn = lambda: ffi.plonk_quotient_fr_device(p[0:5], p[5:14], p[14], p[15], p[16:21], p[21], p[22], p[23], p[24:28], k, sc[0], sc[1], sc[2], sc[3], sc[4], zh, m, factor, out.data_ptr())
    ms = timeit(fn)
    print(f"quotient map n=2^{logn} m={m}: {ms*1e3:9.1f} us  {m/ms/1e6:7.2f} G points/s  {32*33*m/ms/1e6:8.1f} GB/s  ~{95*m/ms/1e6:6.1f} G mul/s ({95*m/ms*1e3/peak*100:.0f}% of mul peak)", flush=True)
    del arrs, out

for logn in (14, 22):
    n = 1 << logn
    c = d(B.random_fr(n, 5)); q = torch.empty(4 * n, dtype=torch.int64, device=dev); v = torch.empty(4, dtype=torch.int64, device=dev)
    z = B.random_fr(1, 6)[0]
    ms = timeit(lambda: ffi.poly_horner_fr_device(c.data_ptr(), n, z, 0, v.data_ptr()))
    print(f"poly eval 2^{logn}: {ms*1e3:9.1f} us  {32*n/ms/1e6:8.1f} GB/s", flush=True)
    ms = timeit(lambda: ffi.poly_horner_fr_device(c.data_ptr(), n, z, q.data_ptr(), v.data_ptr()))
    print(f"poly div (X - z) 2^{logn}: {ms*1e3:9.1f} us  {96*n/ms/1e6:8.1f} GB/s", flush=True)
