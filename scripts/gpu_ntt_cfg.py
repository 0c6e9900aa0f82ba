"""NTT plan-shape sweep (development aid)."""
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
for lg in (12, 14, 16, 18, 20, 22, 24):
    n = 1 << lg
    x = torch.from_numpy(B.random_fr(n, 1).view(np.int64)).to(dev)
    o = torch.empty_like(x); s = torch.empty_like(x)
    for (tile, maxr, two) in ((12, 11, 22), (10, 10, 18), (10, 10, 20), (9, 9, 18), (9, 9, 16), (8, 8, 16), (10, 7, 14), (11, 10, 20)):
        try:
            ffi.configure("ntt_log_tile", tile); ffi.configure("ntt_max_log_r", maxr); ffi.configure("ntt_two_pass_max", two)
            ms = timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n))
            ffi.profile_enable(True); timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n), 5, 0); p = ffi.profile_read("ntt"); ffi.profile_enable(False)
            print(f"ntt 2^{lg} tile=2^{tile} max_r={maxr} two_pass_max={two}: {ms*1e3:8.1f} us  {({k: round(v*1e3,1) for k,v in p['ms'].items()})}", flush=True)
        except Exception as e:
            print(f"ntt 2^{lg} tile=2^{tile} max_r={maxr} two_pass_max={two}: {e}", flush=True)
