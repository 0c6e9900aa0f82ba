"""A/B of the radix-4 stage walk (uzkge_cuda_configure("ntt_radix4", 1)): correctness against the oracle, then timing per size."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench as B
from oracle import cpu as oc  # checker only
from uzkge_b200 import ffi

ffi.init(0)
dev = torch.device("cuda", 0)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ffi.configure("ntt_radix4", 1)       # every size
k = oc.random_fr(1, 31)[0]
bad = 0
for n in [1 << l for l in range(1, 21)] + [3 << l for l in range(1, 19)]:
    x = oc.random_fr(n, 100 + n % 97)
    for inv, cs in ((False, None), (True, None), (False, k), (True, k)):
        got = ffi.ntt_fr(x, n, inv, cs)
        want = oc.ntt_fr(x, n, inverse=inv, coset=cs)
        if not np.array_equal(got, want):
            bad += 1
            print("MISMATCH", n, inv, cs is not None, flush=True)
            break
print("radix-4 correctness:", "ok" if bad == 0 else f"{bad} sizes differ", flush=True)
for lg in (14, 16, 18, 20, 22, 24):
    n = 1 << lg
    x = torch.from_numpy(B.random_fr(n, 1).view(np.int64)).to(dev)
    o, s = torch.empty_like(x), torch.empty_like(x)
    res = {}
    for r4 in (0, 1, 0, 1):
        ffi.configure("ntt_radix4", 1 if r4 else 0)
        res.setdefault(r4, []).append(timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n)))
    print(f"2^{lg}: radix-2 {min(res[0]) * 1e3:8.1f} us   radix-4 {min(res[1]) * 1e3:8.1f} us", flush=True)
n = 3 << 21
x = torch.from_numpy(B.random_fr(n, 1).view(np.int64)).to(dev)
o, s = torch.empty_like(x), torch.empty_like(x)
for r4 in (0, 1):
    ffi.configure("ntt_radix4", r4)
    print(f"3*2^21 radix4={r4}: {timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n)) * 1e3:8.1f} us", flush=True)
sys.exit(1 if bad else 0)
