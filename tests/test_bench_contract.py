"""CPU: the JSON line of `bench.py` as committed under profiles/ (the last run on a B200 of this round) carries every key the
driver's contract names, with consistent values -- a guard against edits to bench.py that drop a field.  The reference arm is run
here for real (it is the CPU arm) on a tiny sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _load(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def _check_line(d, n_gpus):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "roofline", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == n_gpus and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["metric"] == "bn254_g1_msm_2^20_points_per_s" and d["unit"] == "points/s"
    assert abs(d["value"] - n_gpus * (1 << 20) / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 32 << 20 and e["d2h_bytes_per_step"] == 96 and 0 < e["value"] <= d["value"] * 1.05
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "traffic" in r
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_bench_line_follows_the_contract():
    d = _load("r1k_bench.json")
    _check_line(d, 1)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "points/s" and cb["sample"]
    assert d["int_roofline"]["frac"] > 0.85                       # the accumulate kernel against the measured multiplier roof
    circuits = {s.get("circuit", f"synthetic-{s['log_n']}-{s['witness']}-{s['feature_set']}"): s for s in d["plonk"]["sizes"]}
    assert {"zshuffle-52", "zmatchmaking"} <= set(circuits) and "app_errors" not in d["plonk"]
    for s in circuits.values():
        assert s["deterministic"] and s["proof_bytes"] == (1632 if s["feature_set"] == "shuffle" else 1312)
    for name, n in (("r1k_bench_2gpu.json", 2), ("r1k_bench_4gpu.json", 4)):
        _check_line(_load(name), n)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "ntt"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-400:]
    assert len(out.stdout.strip().splitlines()) == 1, "stdout must carry exactly the JSON line"
    d = json.loads(out.stdout)
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm exits non-zero, says why, and prints no result line."""
    from uzkge_b200 import ffi

    if ffi.lib().uzkge_cuda_device_count() > 0:
        import pytest

        pytest.skip("a CUDA device is visible")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and out.stdout.strip() == "" and "no CPU path" in out.stderr


def test_round2_lines_carry_the_binding_roof_the_cpu_prover_and_the_device_group():
    """profiles/r2b_*: the roofline record is on the multiplier pipe with a peak that does not come from the library; the proofs/s
    baseline is a complete CPU proof with the GPU's bytes; the two arms share one config; the multi-GPU lines hold the strong-scaling
    blocks and ONE proof on the device group, all with parity flags set."""
    d = _load("r2b_bench.json")
    r = d["roofline"]
    assert r["bound"] == "imad" and r["unit"] == "T IMAD/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "pipe_probe" in r["peak_source"] and 0.85 < r["frac"] < 1.0 and abs(r["peak"] - r["peak_issue_model"]) / r["peak_issue_model"] < 0.05
    assert "r2b_msm_raw.csv" in r["traffic_source"] and r["traffic"] > 1e9
    assert d["hbm_roofline"]["bound"] == "hbm" and d["hbm_roofline"]["frac"] < 0.05
    assert d["cpu_baseline"]["sample"].startswith("one full 2^20")
    pb = d["plonk"]["cpu_baseline"]
    assert pb["unit"] == "proofs/s" and "proof bytes equal to the GPU's" in pb["sample"] and pb["prove_ms"] > 10 * pb["gpu_prove_ms"]
    for s in d["plonk"]["sizes"]:
        assert s["prover"].startswith("uzkge_cuda_plonk_prove") and s["ops_per_proof"]["msm"] in (13, 16)
    ref = json.loads(open(os.path.join(ROOT, "profiles", "r2b_bench_reference_arm.json")).read().strip().splitlines()[0])
    assert ref["impl"] == "reference" and ref["config"] == d["config"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    one_gpu = [s for s in d["plonk"]["sizes"] if s["log_n"] == 22 and s["witness"] == "uniform"][0]["prove_ms"]
    last = one_gpu
    for n in (2, 4, 8):
        m = _load(f"r2b_bench_{n}gpu.json")
        assert m["n_gpus"] == n and abs(m["value"] - n * (1 << 20) / (m["ms_per_step"] * 1e-3)) / m["value"] < 1e-6
        for key in ("msm_strong", "ntt_strong", "ntt_distributed", "ntt_group_e2e", "commit_group_e2e"):
            assert m[key]["parity_ok"] is True, (n, key)
        assert m["msm_strong"]["ms_per_step"] < m["msm_strong"]["single_gpu_ms"] and m["ntt_strong"]["ms_per_step"] < m["ntt_strong"]["single_gpu_ms"]
        group = [s for s in m["plonk"]["sizes"] if s["mode"] == "device_group"][0]
        split = [s for s in m["plonk"]["sizes"] if s["mode"] == "msm_split"][0]
        assert group["log_n"] == 22 and group["prove_ms"] < split["prove_ms"] < one_gpu and group["prove_ms"] < last
        last = group["prove_ms"]
