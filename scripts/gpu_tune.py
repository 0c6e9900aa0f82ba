"""Device-resident timing sweeps (development aid): python scripts/gpu_tune.py [msm|ntt|all]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
import bench as B

ffi.init(0)
dev = torch.device("cuda", 0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if what in ("ntt", "all"):
    for lg in (14, 16, 18, 20, 22, 24):
        n = 1 << lg
        x = torch.from_numpy(B.random_fr(n, 1).view(np.int64)).to(dev)
        o = torch.empty_like(x); s = torch.empty_like(x)
        for nt in (512, 1024):
            ffi.configure("ntt_big_threads", nt)
            ms = timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n))
            ffi.profile_enable(True); timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n), 5, 0); p = ffi.profile_read("ntt"); ffi.profile_enable(False)
            print(f"ntt 2^{lg} nt={nt}: {ms*1e3:8.1f} us  {n/ms/1e6:8.1f} Gel/s  phases {({k: round(v*1e3,1) for k,v in p['ms'].items()})}", flush=True)
    for n in (49152, 98304, 3 << 21):
        x = torch.from_numpy(B.random_fr(n, 1).view(np.int64)).to(dev)
        o = torch.empty_like(x); s = torch.empty_like(x)
        ms = timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n))
        print(f"ntt mixed {n}: {ms*1e3:8.1f} us  {n/ms/1e6:8.1f} Gel/s", flush=True)

if what in ("msm", "all"):
    tau = B.random_fr(1, 5)[0]
    cfgs = ((14, (0, 11, 12, 13, 14)), (16, (0, 13, 15)), (20, (0, 16, 17, 19, 20)), (22, (0,)))
    if "csweep" in sys.argv:
        cfgs = ((12, (9, 10, 11, 12)), (14, (11, 12, 13, 14)), (16, (13, 14, 15, 16)), (18, (15, 16, 17, 18)), (20, (17, 18, 19, 20, 21)), (22, (19, 20, 21, 22)), (24, (21, 22, 23)))
    elif "quick" in sys.argv:
        cfgs = ((14, (13,)), (16, (15,)), (20, (17, 20)), (22, (20,)))
    for lg, cs in cfgs:
        n = 1 << lg
        bases = ffi.srs_generate(tau, n)
        sc = torch.from_numpy(B.random_fr(n, 2).view(np.int64)).to(dev)
        out = torch.zeros(12, dtype=torch.int64, device=dev)
        for c in cs:
            h = ffi.srs_upload(bases, c)
            info = ffi.srs_info(h)
            for lanes in ((0,) if "csweep" in sys.argv else (0, -1) if "quick" in sys.argv else (0,) if lg >= 20 and c not in (0,17) else (0, 1, 2, 4, 8, 16, 32)):
                if lanes < 0: continue
                ffi.configure("msm_lanes", lanes)
                ms = timeit(lambda: ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr()), 5, 2)
                ffi.profile_enable(True); timeit(lambda: ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr()), 3, 0); p = ffi.profile_read("msm"); ffi.profile_enable(False)
                print(f"msm 2^{lg} c={info['window_bits']} W={info['windows']} lanes={lanes}: {ms*1e3:8.1f} us  {n/ms/1e3:8.1f} Mpts/s  {({k: round(v*1e3,1) for k,v in p['ms'].items()})}", flush=True)
            ffi.configure("msm_lanes", 0)
            ffi.srs_free(h)

if what in ("batch",):
    tau = B.random_fr(1, 5)[0]
    for lg in (12, 13, 14, 16):
        n = 1 << lg
        bases = ffi.srs_generate(tau, n)
        h = ffi.srs_upload(bases, 0)
        info = ffi.srs_info(h)
        K = 16
        scs = [torch.from_numpy(B.random_fr(n, 20 + j).view(np.int64)).to(dev) for j in range(K)]
        out = torch.zeros(12 * K, dtype=torch.int64, device=dev)
        for k in (1, 2, 4, 8, 16):
            ptrs = [t.data_ptr() for t in scs[:k]]
            ms = timeit(lambda: ffi.msm_g1_batch_device(h, ptrs, [n] * k, out.data_ptr()), 5, 2)
            ffi.profile_enable(True); timeit(lambda: ffi.msm_g1_batch_device(h, ptrs, [n] * k, out.data_ptr()), 3, 0); p = ffi.profile_read("msm"); ffi.profile_enable(False)
            print(f"msm batch 2^{lg} c={info['window_bits']} slots={info['batch_slots']} k={k}: {ms*1e3:8.1f} us total {ms*1e3/k:8.1f} us/MSM  {({kk: round(v*1e3,1) for kk,v in p['ms'].items()})}", flush=True)
        ffi.srs_free(h)
