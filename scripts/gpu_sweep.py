"""BASELINE.json configs[1] and configs[2] sweeps, device-resident (CUDA events): G1 MSM 2^16..2^24 and Fr NTT / iNTT /
coset-FFT 2^16..2^24 plus the mixed-radix sizes 3*2^k the quotient domain needs.  Writes one JSON document.
Every row is CHECKED before it is written (a timing of an unverified result is refused): MSM results against the SRS trapdoor
(MSM(srs, f) = f(tau) G, CPU oracle: Horner + one scalar multiplication); transforms through Horner spot checks of the outputs at
random indices on the CPU oracle, and the inverse variants through the round trip.

    python scripts/gpu_sweep.py [out.json] [max_log_msm]
"""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi
from oracle import cpu as oc   # checker only
import bench as B

ffi.init(0)
dev = torch.device("cuda", 0)
out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.json"
max_log_msm = int(sys.argv[2]) if len(sys.argv) > 2 else 24


def timeit(fn, reps=10, warm=3):
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


peak = ffi.bench_field_mul("fq", 2000)
doc = {"fq_mul_peak_per_s": peak, "msm": [], "ntt": []}

k_shift = B.random_fr(1, 77)[0]
for lg, mixed in [(l, False) for l in range(16, 25)] + [(l, True) for l in range(15, 24)]:
    n = (3 << lg) if mixed else (1 << lg)
    x = torch.from_numpy(B.random_fr(n, 1).view(np.int64)).to(dev)
    o = torch.empty_like(x)
    s = torch.empty_like(x)
    row = {"n": n, "label": ("3*2^%d" % lg) if mixed else ("2^%d" % lg)}
    # checks first: fft / coset fft outputs at two random indices equal the polynomial's value there (Horner on the CPU oracle);
    # ifft(fft(x)) = x and coset_ifft(coset_fft(x)) = x with the inverse shift
    xh = x.cpu().numpy().view(np.uint64).reshape(n, 4)
    w = ffi.fr_root_of_unity(n)
    k_inv = oc.fr_inv(k_shift)
    for cs_ in (None, k_shift):
        ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n, False, cs_)
        oh = o.cpu().numpy().view(np.uint64).reshape(n, 4)
        for idx in (1, int(np.random.default_rng(lg).integers(2, n))):
            point = oc.fr_pow(w, idx) if cs_ is None else oc.fr_mul(oc.fr_pow(w, idx).reshape(1, 4), cs_.reshape(1, 4))[0]
            if not np.array_equal(oc.fr_eval(xh, point), oh[idx]):
                raise SystemExit(f"sweep: transform {row['label']} differs from the oracle at index {idx} -- refusing to write the row")
        ffi.ntt_fr_device(o.data_ptr(), o.data_ptr(), s.data_ptr(), n, n, True, None if cs_ is None else k_inv)
        if not torch.equal(o, x):
            raise SystemExit(f"sweep: inverse transform {row['label']} does not return the input -- refusing to write the row")
    row["checked"] = "Horner spot checks (2 indices, fft and coset fft) + inverse round trips"
    for name, inv, cs in (("fft", False, None), ("ifft", True, None), ("coset_fft", False, k_shift), ("coset_ifft", True, k_shift)):
        ms = timeit(lambda: ffi.ntt_fr_device(x.data_ptr(), o.data_ptr(), s.data_ptr(), n, n, inv, cs), 10 if n < (1 << 23) else 5)
        row[name + "_us"] = round(ms * 1e3, 1)
    row["fft_elements_per_s"] = n / row["fft_us"] * 1e6
    doc["ntt"].append(row)
    print(row, flush=True)
    del x, o, s

tau = B.random_fr(1, 5)[0]
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
SMALL = np.array([[((v << 256) % FR) >> (64 * j) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for v in range(1 << 16)], dtype=np.uint64)
for lg in range(16, max_log_msm + 1):
    n = 1 << lg
    bases = ffi.srs_generate(tau, n)
    h = ffi.srs_upload(bases, 0)
    info = ffi.srs_info(h)
    row = {"n": n, "label": "2^%d" % lg, "window_bits": info["window_bits"], "windows": info["windows"],
           "device_bytes": info["device_bytes"], "precompute_ms": info["precompute_ms"]}
    out = torch.zeros(12, dtype=torch.int64, device=dev)
    for kind, seed in (("uniform", 2), ("witness", 3)):
        sc_h = B.random_fr(n, seed)
        if kind == "witness":  # 50 % zero, 30 % in {0,1}, 10 % < 2^16, 10 % uniform
            rng = np.random.default_rng(9)
            u = rng.random(n)
            sc_h = sc_h.copy()
            sc_h[u < 0.5] = 0
            m = (u >= 0.5) & (u < 0.8)
            sc_h[m] = SMALL[rng.integers(0, 2, int(m.sum()))]
            m = (u >= 0.8) & (u < 0.9)
            sc_h[m] = SMALL[rng.integers(0, 1 << 16, int(m.sum()))]
        sc = torch.from_numpy(sc_h.view(np.int64)).to(dev)
        ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr())
        got = out.cpu().numpy().view(np.uint64)
        want = oc.g1_mul(bases[0], oc.fr_eval(sc_h, tau))
        if not np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(want)):
            raise SystemExit(f"sweep: MSM 2^{lg} ({kind}) differs from f(tau) G -- refusing to write the row")
        row["checked"] = "MSM(srs, f) == f(tau) G on the CPU oracle, uniform and witness-like scalars"
        ms = timeit(lambda: ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr()), 5, 2)
        row[kind + "_us"] = round(ms * 1e3, 1)
        row[kind + "_points_per_s"] = n / ms * 1e3
        if kind == "uniform":
            ffi.profile_enable(True)
            timeit(lambda: ffi.msm_g1_device(h, sc.data_ptr(), n, out.data_ptr()), 3, 0)
            p = ffi.profile_read("msm")
            ffi.profile_enable(False)
            row["phases_us"] = {k: round(v * 1e3, 1) for k, v in p["ms"].items()}
        del sc
    doc["msm"].append(row)
    print(row, flush=True)
    ffi.srs_free(h)
    del bases

os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
with open(out_path, "w") as f:
    json.dump(doc, f, indent=1)
print("wrote", out_path)
