"""Quick GPU bring-up check (development aid): field mul, NTT and MSM through the C ABI against the CPU oracle."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from uzkge_b200 import ffi
from oracle import cpu as oc

def t(name, ok, extra=""):
    print(("PASS " if ok else "FAIL ") + name, extra, flush=True)
    return ok

ffi.init(0)
print(ffi.version())
allok = True
a = oc.random_fr(5000, 1); b = oc.random_fr(5000, 2)
allok &= t("fr_mul", np.array_equal(ffi.field_mul(a, b, "fr"), oc.fr_mul(a, b)))
allok &= t("fq_mul", np.array_equal(ffi.field_mul(a, b, "fq"), oc.fq_mul(a, b)))
for f in ("fr", "fq"):
    print("bench", f, "%.1f Gmul/s" % (ffi.bench_field_mul(f, 2000) / 1e9), flush=True)

sizes = [1, 2, 4, 8, 16, 64, 1024, 2048, 4096, 8192, 1 << 14, 1 << 16, 3, 6, 12, 48, 96, 3 << 10, 3 << 12, 49152, 98304, 1 << 20, 3 << 19]
if "--big" in sys.argv:
    sizes += [1 << 22, 1 << 23, 1 << 24, 3 << 21]
k = oc.random_fr(1, 77)[0]
for n in sizes:
    x = oc.random_fr(n, 100 + n % 97)
    for inv in (False, True):
        for cs in (None, k):
            for len_in in sorted({n, max(1, n // 2 + 1)}):
                t0 = time.time()
                got = ffi.ntt_fr(x[:len_in], n, inv, cs)
                t1 = time.time()
                want = oc.ntt_fr(x[:len_in], n, inv, cs)
                ok = np.array_equal(got, want)
                allok &= t(f"ntt n={n} inv={inv} coset={cs is not None} len_in={len_in}", ok, "%.1f ms" % ((t1 - t0) * 1e3))
                if not ok:
                    bad = np.nonzero((got != want).any(axis=1))[0]
                    print("   mismatches:", bad.size, "first", bad[:8])

for n, cbits in [(1, 0), (2, 0), (33, 0), (1000, 0), (1000, 4), (1000, 9), (4096, 0), (16384, 0), (16384, 16), (1 << 16, 0), (1 << 18, 0)]:
    pts = oc.g1_random_points(n, 5)
    sc = oc.random_fr(n, 6)
    t0 = time.time(); h = ffi.srs_upload(pts, cbits); t1 = time.time()
    info = ffi.srs_info(h)
    for lanes in (0, 1, 32):
        ffi.configure("msm_lanes", lanes)
        t2 = time.time(); got = ffi.msm_g1(h, sc); t3 = time.time()
        want = oc.msm_g1(pts, sc)
        ok = np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(want))
        allok &= t(f"msm n={n} c={info['window_bits']} lanes={lanes}", ok, "upload %.1f ms (pre %.1f) msm %.2f ms" % ((t1 - t0) * 1e3, info["precompute_ms"], (t3 - t2) * 1e3))
    ffi.configure("msm_lanes", 0)
    # skewed scalars
    sk = sc.copy(); sk[: n // 2] = 0; one = oc.fr_to_mont(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]; sk[n // 2 : (3 * n) // 4] = one
    got = ffi.msm_g1(h, sk); want = oc.msm_g1(pts, sk)
    allok &= t(f"msm skew n={n}", np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(want)))
    ffi.srs_free(h)
print("ALL OK" if allok else "SOME FAILED")
sys.exit(0 if allok else 1)
