"""A/B of the bucket accumulation: XYZZ kernel vs batched-affine tree (uzkge_cuda_configure "msm_affine")."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
import bench as B
ffi.init(0)
dev = torch.device("cuda", 0)
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tau = B.random_fr(1, 5)[0]
for lg in [int(a) for a in sys.argv[1:]] or [18, 20, 22]:
    n = 1 << lg
    bases = ffi.srs_generate(tau, n)
    h = ffi.srs_upload(bases, 0)
    info = ffi.srs_info(h)
    sc = torch.from_numpy(B.random_fr(n, 2).view(np.int64)).to(dev)
    outs = {}
    for mode in (0, 1, 0, 1):
        ffi.configure("msm_affine", mode)
        out = torch.zeros(12, dtype=torch.int64, device=dev)
        ms = timeit(lambda: ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr()))
        ffi.profile_enable(True); timeit(lambda: ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr()), 3, 0); p = ffi.profile_read("msm"); ffi.profile_enable(False)
        outs[mode] = ffi.g1_to_affine(out.cpu().numpy().view(np.uint64))
        print(f"msm 2^{lg} c={info['window_bits']} affine={mode}: {ms*1e3:8.1f} us  accumulate {p['ms']['accumulate']*1e3:8.1f} us  reduce {p['ms']['reduce']*1e3:6.1f}", flush=True)
    print("  same result:", bool(np.array_equal(outs[0], outs[1])), flush=True)
    ffi.configure("msm_affine", 0)
    ffi.srs_free(h)
