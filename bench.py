#!/usr/bin/env python
"""bench.py -- BN254 G1 MSM 2^20 points/s (headline, BASELINE.json configs[1]), Fr NTT 2^22 elements/s (configs[2]) and
TurboPlonK proofs/s on synthetic circuits (configs[4]) on N B200s, one process per GPU, next to the CPU restatement of the
reference's arkworks path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|both|msm|ntt|plonk]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch of synthetic input:
  msm : one variable-base MSM over 2^20 resident KZG bases (powers-of-tau SRS) with fresh uniform Fr scalars.  The K timed
        steps travel in ONE batch call, which the engine software-pipelines (`detail.single_call_ms`: one MSM per call).
        N > 1: the job is ONE MSM over N * 2^20 points; rank r owns slice r (bases resident on its GPU) and the
        N partial sums (96 B each) are all-gathered and added -- weak scaling, no other collective.
  ntt : one forward 2^22 Fr NTT, natural order in and out (block "ntt"; replicas at N > 1: it fits one GPU).
  plonk : one complete proof of a synthetic TurboPlonK circuit (block "plonk": 2^14 and 2^22 gates, default and shuffle feature
        sets, uniform and bits witnesses, the zshuffle-52 and zmatchmaking circuits; compiled prover behind the C ABI, the Python
        mirror timed beside it; `cpu_baseline`: a complete CPU proof of the same statement, bytes compared).  N > 1: replicas for
        throughput, and ONE 2^22 proof three ways: "device_group" (one process drives all N GPUs through one C-ABI call per
        proof), "msm_split" (round 1's Python split prover, one process per GPU), "replicas".
  strong scaling (N > 1, blocks "msm_strong", "ntt_strong", "ntt_distributed"): ONE 2^20 MSM and ONE 2^22 / 2^24 transform over the
        N GPUs, each with the single-GPU time of the same run, the efficiency and the bytes exchanged; the transforms with cyclic
        and natural output, over peer memory (exchanges fused into the kernels) and over NCCL all-to-alls.
The JSON line's top level is the MSM (`value` device-resident with CUDA events on the launching stream, max over ranks; `e2e`
through the host-pointer C ABI a Rust caller uses: pinned host buffers, H2D + D2H inside the timed region; `roofline`,
`int_roofline`, `cpu_baseline`, `clocks`, `gpu_launches`).  Inputs rotate over more distinct buffers than fit in the 126 MB L2
(`config.l2`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
LOG_MSM = 20
LOG_NTT = 22


# ------------------------------------------------------------------------------------------------ inputs

_JSON_OUT = None      # the process's real stdout once __main__ has redirected file descriptor 1 (see the bottom of the file)


def emit(line: dict) -> None:
    """Print the result line on the real stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def random_fr(n: int, seed: int) -> np.ndarray:
    """n uniform Fr elements as raw Montgomery limbs (uniform residues stay uniform under the Montgomery map)."""
    rng = np.random.default_rng(seed)
    mod = np.array([(FR >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)

    def draw(k):
        a = rng.integers(0, 1 << 63, size=(k, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(k, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        return a

    def ge_mod(a):
        ge = np.zeros(a.shape[0], dtype=bool)
        und = np.ones(a.shape[0], dtype=bool)
        for k in (3, 2, 1, 0):
            gt = und & (a[:, k] > mod[k])
            lt = und & (a[:, k] < mod[k])
            ge |= gt
            und &= ~(gt | lt)
        return ge | und

    a = draw(n)
    while True:
        bad = np.nonzero(ge_mod(a))[0]
        if bad.size == 0:
            return a
        a[bad] = draw(bad.size)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self._t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self._t = threading.Thread(target=self._read, daemon=True)
        self._t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, smax, reasons = [], None, set()
        for r in inside:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}



# ------------------------------------------------------------------------------------------------ roofline inputs
# ncu `--set full` captures of the CURRENT kernels, exported with `--page raw --csv` (profiles/README.md says how each was taken)
TRAFFIC_FILES = {"msm": ["profiles/r2b_msm_raw.csv", "profiles/r2_msm_accumulate_raw.csv", "profiles/r1c_msm_raw.csv"],
                 "ntt": ["profiles/r2b_ntt_raw.csv", "profiles/r2_ntt_raw.csv", "profiles/r1c_ntt_raw.csv"]}
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_traffic(kind: str, kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch whose name contains `kernel_substr`, read from the committed
    ncu raw page (per launch).  Returns (bytes or None, source description)."""
    import csv

    for rel in TRAFFIC_FILES[kind]:
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        try:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            for r in rows[2:]:
                if kernel_substr in r[ik]:
                    total = float(r[ir].replace(",", "")) * _UNIT[units[ir]] + float(r[iw].replace(",", "")) * _UNIT[units[iw]]
                    return total, f"{rel}: dram__bytes_read.sum + dram__bytes_write.sum of `{r[ik]}` (ncu --set full, one launch)"
        except Exception as exc:  # a malformed capture must not break the bench line
            return None, f"{rel}: unreadable ({exc!r})"
    return None, "no ncu capture committed for this kernel"


def imad_probe(gpu_index: int) -> dict:
    """The integer roof from a program that shares no code with the library (scripts/cuda/pipe_probe.cu: dependent IMAD.WIDE chains
    on every SM, CUDA events).  Returns {} when the binary is missing (build() compiles it)."""
    exe = os.path.join(ROOT, "scripts", "cuda", "pipe_probe")
    if not os.path.exists(exe):
        return {}
    try:
        out = subprocess.run([exe, "--json", str(gpu_index)], capture_output=True, text=True, timeout=120)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception:
        return {}


# multiplier-pipe instructions per field product (ff.cuh; SASS excerpt in profiles/r2_fe_mul_sass.txt): 8 x (8 + 8 + 1) for fe_mul,
# 36 + 8 x 9 for the dedicated fe_sqr
IMAD_PER_MUL, IMAD_PER_SQR = 136, 108
IMAD_PER_MADD = 8 * IMAD_PER_MUL + 2 * IMAD_PER_SQR      # XYZZ mixed addition: 8 M + 2 S

# ------------------------------------------------------------------------------------------------ reference arm
def headline_config(head: str, world: int) -> dict:
    """The `config` of the JSON line -- the SAME dictionary in both arms (`--impl ours` and `--impl reference`), so that the driver
    compares like with like; what is specific to one arm lives in that arm's `detail` / `reference_note`."""
    return {
        "workload": ("BN254 G1 variable-base MSM, 2^20 points per GPU, bases resident (affine KZG SRS), uniform Fr scalars"
                     if head == "msm" else "BN254 Fr radix-2 NTT 2^22, natural order in/out"),
        "parallelism": f"{world} process(es), one per GPU; MSM points split per GPU, partial sums all-gathered (96 B) and added",
        "pipelining": "GPU arm: the K steps are submitted in one batch call and software-pipelined by the engine (sort / reduce of "
                      "neighbouring steps under the accumulate kernel); detail.single_call_ms is one MSM per call",
        "l2": "GPU arm: inputs rotate over 8 x 32 MiB scalar sets / 4 x 128 MiB vectors (> 126 MB L2); the MSM's window tables are 0.8 GiB",
    }


def run_reference(args, rank: int, world: int) -> int:
    """The reference's CPU path for the same metric: no Rust toolchain exists here or on the GPU box and the
    arithmetic lives in un-vendored arkworks forks, so this arm times the oracle port
    (oracle/oracle.c: arkworks' signed-digit Pippenger / radix-2 FFT restated in C + OpenMP) on all host cores."""
    if rank != 0:
        return 0
    from oracle import cpu as oc

    oc.set_num_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: use every host core we may use
    cores = oc.num_threads()
    log_n = LOG_MSM if args.workload not in ("ntt", "plonk") else (LOG_NTT if args.workload == "ntt" else int(args.plonk_logs.split(",")[0]))
    n = 1 << log_n
    units = n
    t_setup = time.time()
    if args.workload == "plonk":
        # a complete CPU proof per step: the compiled CPU prover (oracle/cpu_prover.py + oracle.c) on a synthetic circuit of add / mul
        # gates in 8 layers (the shape of plonk.TurboCS.synthetic), SRS = random curve points (the prover never needs the trapdoor)
        from oracle import cpu_prover as cpp
        from oracle import plonk_prover as opp

        rnd = np.random.default_rng(0xB2000004)
        sel = np.zeros((9, n), dtype=np.int64)
        wir = np.zeros((5, n), dtype=np.int64)
        sel[6, 1] = 1
        sel[8, :2] = 1
        wir[:, 1] = 1
        g = n - 2
        per = [g // 8 + (1 if i < g % 8 else 0) for i in range(8)]
        n_inputs = max(per[0], 2)
        wit = [0, 1] + [int(x) for x in rnd.integers(1, 1 << 62, n_inputs)]
        prev_lo, prev_n, row = 2, n_inputs, 2
        for cnt in per:
            a_, b_ = prev_lo + rnd.integers(0, prev_n, cnt), prev_lo + rnd.integers(0, prev_n, cnt)
            mul = rnd.integers(0, 2, cnt).astype(bool)
            var = len(wit)
            wir[0, row:row + cnt], wir[1, row:row + cnt], wir[4, row:row + cnt] = a_, b_, np.arange(var, var + cnt)
            sel[0, row:row + cnt] = sel[1, row:row + cnt] = ~mul
            sel[4, row:row + cnt] = mul
            sel[8, row:row + cnt] = 1
            wit += [(wit[x] * wit[y] if m_ else wit[x] + wit[y]) % FR for x, y, m_ in zip(a_.tolist(), b_.tolist(), mul.tolist())]
            prev_lo, prev_n, row = var, cnt, row + cnt
        table = cpp.A([0, 1])
        acs = cpp.ArrayCS([table[sel[i]] for i in range(9)], wir)
        cpcs = cpp.CpuKzg(oc.g1_random_points(n + 3, 0xB2000006))
        cP = cpp.indexer(acs, cpcs)
        wit_arr = cpp.A(wit)
        units = 1

        def step(i):
            cpp.prover(opp.ChaCha(bytes(32)), opp.Transcript(b"bench"), cpcs, acs, cP, wit_arr)
        metric, unit, wl = "turboplonk_synthetic_proofs_per_s", "proofs/s", f"synthetic TurboPlonK circuit, 2^{log_n} gates, full prove"
    elif args.workload != "ntt":
        pts = oc.g1_random_points(n, 0xB2000001)
        scal = [oc.random_fr(n, 0xB2000002 + i) for i in range(2)]

        def step(i):
            oc.msm_g1(pts, scal[i % 2])
        metric, unit, wl = "bn254_g1_msm_2^20_points_per_s", "points/s", f"G1 MSM 2^{log_n}"
    else:
        vec = [oc.random_fr(n, 0xB2000003 + i) for i in range(2)]

        def step(i):
            oc.ntt_fr(vec[i % 2], n)
        metric, unit, wl = "bn254_fr_ntt_2^22_elements_per_s", "elements/s", f"Fr NTT 2^{log_n}"
    for i in range(args.warmup):
        step(i)
    t0 = time.time()
    for i in range(args.steps):
        step(i)
    dt = time.time() - t0
    v = units * args.steps / dt
    head = "msm" if args.workload not in ("ntt", "plonk") else args.workload
    line = {
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 limbs (256-bit Montgomery)", "data": "synthetic",
        "config": headline_config(head, args.gpus) if head != "plonk" else {"workload": wl + ", witness resident in HBM"},
        "reference_note": {"workload": wl, "where": "host CPU", "note": "oracle port of arkworks' algorithms (no Rust toolchain: the "
                           "reference itself cannot be built); full workload per step; MSM bases are random curve points in host "
                           "memory (the GPU arm's are a powers-of-tau SRS: the cost of an MSM does not depend on which points)",
                           "setup_s": round(t0 - t_setup, 1)},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": f"{args.steps} x full {wl}"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------ PlonK proofs
def plonk_proofs(args, dev, K: int, W: int, rank: int = 0, world: int = 1, barrier=None, max_over_ranks=None) -> dict:
    """BASELINE.json's first metric: full TurboPlonK proofs/s on synthetic circuits (configs[4]; 2^14 is the size of the zshuffle
    circuit, configs[0]).  A step = one complete `prover` call (5 rounds, 13 MSMs, 7 iFFT(n) + 7 coset FFT(6n) + 1 coset iFFT(6n),
    quotient map, openings) -- wall clock around a synchronised call, because the Fiat-Shamir transcript puts the host in the loop.
    `proofs_per_s`: witness resident in HBM; `e2e_proofs_per_s`: witness in host memory, uploaded inside the timed region.
    N > 1 GPUs: circuits up to 2^18 gates run as N independent provers (replicas; proofs are latency-bound there); larger ones
    as ONE proof driven by rank 0 (dist.SplitCommitter): every commitment's points are split over the N GPUs and the six
    independent 6n transforms of the quotient round run one per GPU."""
    import torch

    from uzkge_b200 import KZGCommitmentSchemeBN254, ffi, plonk
    from uzkge_b200 import dist as udist
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    tau = plonk.mont(0x1234567890ABCDEF1234567890ABCDEF)
    out = {"metric": "turboplonk_synthetic_proofs_per_s", "unit": "proofs/s", "sizes": []}
    logs = [int(x) for x in args.plonk_logs.split(",") if x]
    # (log size, witness, shuffle feature set, one proof split over the GPUs)
    runs = [(lg, "uniform", False, world > 1 and lg > 18) for lg in logs]
    if logs and world > 1 and logs[-1] > 18:
        # ONE proof on all N GPUs driven by ONE process through the C ABI (uzkge_cuda_plonk_params_upload_multi): rank 0 owns the device
        # group, the other ranks wait on the rendezvous store (no kernel of theirs runs meanwhile)
        runs.insert(len(logs), (logs[-1], "uniform", False, "group"))
    if logs and world == 1:
        if logs[-1] >= 20:
            runs.append((logs[-1], "bits", False, False))    # the same circuit shape over a witness of bits / small integers
        runs.append((logs[0], "uniform", True, False))       # the `shuffle` feature set: the 1632-byte proof format of zshuffle's verifier
    if logs and world > 1 and logs[-1] > 18:
        runs.append((logs[-1], "uniform", False, False))     # throughput: N independent provers of the large circuit (25 GiB each)
    for lg, witness, shuffle_features, split in runs:
        n = 1 << lg
        steps = K if lg <= 18 else max(2, min(K, 5))
        torch.cuda.reset_peak_memory_stats()          # per-size peak (torch allocations); the library's own arenas are in device_used_gib
        t0 = time.perf_counter()
        lagrange = None
        group = split == "group"
        if group:
            import torch.distributed as tdist

            store = tdist.distributed_c10d._get_default_store()
            key = f"uzkge_group_proof_done_{lg}"
            if rank != 0:
                torch.cuda.synchronize()
                store.wait([key], __import__("datetime").timedelta(minutes=30))   # host-side wait: this rank's GPU belongs to rank 0's device group meanwhile
                barrier()
                continue
            split = False
            ffi.configure("virtual_devices", 0)
            if ffi.init_devices(world) != world:
                raise SystemExit("bench.py: the device group does not cover the GPUs of the job")
        if split:
            bases = ffi.srs_generate(tau, n + 3)
            pcs = udist.SplitCommitter(bases, rank, world, device=dev)
            del bases
            if rank != 0:
                pcs.serve()
                pcs.close()
                barrier()
                torch.cuda.empty_cache()
                continue
        else:
            pcs = KZGCommitmentSchemeBN254.new(n + 2, tau)
            lagrange = KZGCommitmentSchemeBN254.new_lagrange(n, tau)   # prover_with_lagrange: what zshuffle / zmatchmaking call
        cs = plonk.TurboCS.synthetic(lg, seed=0xB2000004 + (0 if split else rank), witness=witness)
        params = plonk.indexer(cs, pcs, shuffle=shuffle_features)
        torch.cuda.synchronize()
        setup_s = time.perf_counter() - t0
        wit_pinned = ffi.PinnedArray(cs.get_witness_array().shape)     # the caller's witness, page-locked (uzkge_cuda_host_alloc)
        wit_pinned.array[:] = cs.get_witness_array()
        wit_host = wit_pinned.array
        wit = plonk.DevVec.from_numpy(wit_host, dev)
        # one GPU: the compiled prover behind the C ABI (uzkge_cuda_plonk_prove); the Python mirror is timed beside it.  A proof split
        # over the GPUs of the box is driven by the mirror (dist.SplitCommitter)
        native = None if split else NativeProver(cs, params, pcs, lagrange, multi=group)

        def prove(w, timings=None):
            if native is not None:
                return native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), w, timings=timings)
            return plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, w, timings=timings, lagrange_pcs=lagrange)

        for _ in range(min(W, 3)):
            proof = prove(wit)
        torch.cuda.synchronize()
        if not split and not group and barrier:
            barrier()
        timings = {}
        l0 = ffi.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            prove(wit, timings)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        launches = (ffi.launch_count() - l0) // steps
        if not split and not group and barrier:
            barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            proof2 = prove(wit_host)
        torch.cuda.synchronize()
        dt_e2e = (time.perf_counter() - t0) / steps
        same = all(a == b for a, b in zip(proof.cm_t_vec + [proof.opening_witness_zeta], proof2.cm_t_vec + [proof2.opening_witness_zeta]))
        mirror_ms, ops = None, None
        if native is not None:
            ops = {k: native.last_stats[k] for k in ("msm", "ifft_n", "fft_n", "coset_fft_m", "coset_ifft_m", "evals")}
            ops["quotient_points"] = int(params.m)
            msteps = max(1, min(steps, 5))
            mproof = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit, lagrange_pcs=lagrange)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(msteps):
                plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit, lagrange_pcs=lagrange)
            torch.cuda.synchronize()
            mirror_ms = (time.perf_counter() - t0) / msteps * 1e3
            if mproof.to_bytes_be() != proof.to_bytes_be():
                raise SystemExit("bench.py: the compiled prover's proof differs from the Python mirror's -- refusing to report a number")
            native.close()
        if split:
            pcs.shutdown()
            window_bits = ffi.srs_info(pcs.handle)["window_bits"]
            proofs_in_flight = 1
        elif group:
            window_bits = pcs.info()["window_bits"]
            proofs_in_flight = 1
        else:
            window_bits = pcs.info()["window_bits"]
            proofs_in_flight = world
            if world > 1:
                dt, dt_e2e = max_over_ranks(dt), max_over_ranks(dt_e2e)
        out["sizes"].append({
            "log_n": lg, "n_gpus": world,
            "mode": "msm_split" if split else ("device_group" if group else ("replicas" if world > 1 else "single")),
            "witness": witness, "lagrange_commitments": lagrange is not None,
            "feature_set": "shuffle" if shuffle_features else "default", "proof_bytes": len(proof.to_bytes_be()),
            "prove_ms": dt * 1e3, "proofs_per_s": proofs_in_flight / dt, "e2e_prove_ms": dt_e2e * 1e3,
            "e2e_proofs_per_s": proofs_in_flight / dt_e2e,
            "h2d_bytes_per_step": int(wit_host.nbytes), "d2h_bytes_per_step": 13 * 96 + 16 * 32, "steps": steps,
            "launches_per_proof": int(launches), "rounds_ms": {k: v / steps for k, v in timings.items()},
            "prover": ("uzkge_cuda_plonk_prove over a multi-device parameter handle (one process, one C-ABI call per proof: commitments by SRS "
                       "slice, quotient by cosets over peer memory)" if group else
                       "uzkge_cuda_plonk_prove (compiled, one C-ABI call per proof)" if native is not None else
                       "uzkge_b200.plonk.prover (Python mirror, one C-ABI call per operation)"),
            "python_mirror_prove_ms": mirror_ms,
            "ops_per_proof": ops, "ops_source": "counted by the prover (uzkge_plonk_proof)" if ops else None,
            "setup_s": setup_s, "deterministic": bool(same),
            "window_bits": window_bits, "hbm_peak_gib": torch.cuda.max_memory_allocated() / 2**30,
            "device_used_gib": (lambda fr_, tot_: (tot_ - fr_) / 2**30)(*torch.cuda.mem_get_info()),
        })
        if world == 1 and lg <= 16 and not shuffle_features and "_cpu_inputs" not in out:
            # what the CPU prover of the cpu_baseline leg needs to prove the SAME statement: the circuit as arrays, the witness, the SRS
            # points, and the GPU's proof to hold the CPU's against
            out["_cpu_inputs"] = {
                "log_n": lg, "selectors": np.array(cs.selectors), "wiring": np.array(cs.wiring), "witness": np.array(wit_host),
                "boolean": list(getattr(cs, "boolean_constraint_indices", [])),
                "pub_constraint": list(cs.public_vars_constraint_indices), "pub_witness": list(cs.public_vars_witness_indices),
                "srs": np.array(pcs.public_parameter_group_1), "proof_bytes": proof.to_bytes_be(), "gpu_prove_ms": dt * 1e3,
            }
        pcs.close()
        if lagrange is not None:
            lagrange.close()
        del wit_host
        wit_pinned.free()
        del params, wit, cs, pcs, lagrange
        torch.cuda.empty_cache()
        if group:
            store.set(key, "1")
            barrier()
        if split:
            barrier()
    return out


def _app_circuit_proofs(dev, K: int, W: int, name: str, cs, build_s: float, shuffle: bool, label: bytes, count: int, apk=None) -> dict:
    """One of the reference's application circuits through indexer_with_lagrange / prover_with_lagrange the way the production path
    does: every commitment over the Lagrange SRS (lagrange_all).  The SRS is synthetic (known tau) -- the parity tests use the
    bundled production parameters.  A step = one `prover` call; building the circuit (host, Python integers) is reported beside it."""
    import torch

    from uzkge_b200 import KZGCommitmentSchemeBN254, ffi, plonk
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    tau = plonk.mont(0x1234567890ABCDEF1234567890ABCDEF)
    n = cs.size
    t0 = time.perf_counter()
    pcs, lagrange = KZGCommitmentSchemeBN254.new(n + 2, tau), KZGCommitmentSchemeBN254.new_lagrange(n, tau)
    params = plonk.indexer(cs, pcs, shuffle=shuffle, lagrange_pcs=lagrange)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    refresh_ms = None
    if apk is not None:
        t0 = time.perf_counter()
        plonk.refresh_prover_params_public_key(cs, params, pcs, apk, lagrange_pcs=lagrange)
        refresh_ms = (time.perf_counter() - t0) * 1e3
    wit_pinned = ffi.PinnedArray(cs.get_witness_array().shape)
    wit_pinned.array[:] = cs.get_witness_array()
    wit_host = wit_pinned.array

    from uzkge_b200.native import NativeProver

    native = NativeProver(cs, params, pcs, lagrange, True)

    def prove(w, timings=None):
        tr = Transcript(label)
        tr.append_u64(count)
        return native.prove(ChaChaRng.from_seed(bytes(32)), tr, w, timings=timings)

    def prove_mirror(w):
        tr = Transcript(label)
        tr.append_u64(count)
        return plonk.prover(ChaChaRng.from_seed(bytes(32)), tr, pcs, cs, params, w, lagrange_pcs=lagrange, lagrange_all=True)

    wit = plonk.DevVec.from_numpy(wit_host, dev)
    for _ in range(max(W, 3)):
        proof = prove(wit)
    torch.cuda.synchronize()
    timings = {}
    l0 = ffi.launch_count()
    t0 = time.perf_counter()
    for _ in range(K):
        prove(wit, timings)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    launches = (ffi.launch_count() - l0) // K
    t0 = time.perf_counter()
    for _ in range(K):
        proof2 = prove(wit_host)
    torch.cuda.synchronize()
    dt_e2e = (time.perf_counter() - t0) / K
    n_msm, n_sel = (16, 3) if shuffle else (13, 0)
    ops = {k: native.last_stats[k] for k in ("msm", "ifft_n", "fft_n", "coset_fft_m", "coset_ifft_m", "evals")}
    ops["quotient_points"] = int(params.m)
    mproof = prove_mirror(wit)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(min(K, 10)):
        prove_mirror(wit)
    torch.cuda.synchronize()
    mirror_ms = (time.perf_counter() - t0) / min(K, 10) * 1e3
    if mproof.to_bytes_be() != proof.to_bytes_be():
        raise RuntimeError("the compiled prover's proof differs from the Python mirror's")
    native.close()
    res = {
        "circuit": name, "log_n": n.bit_length() - 1, "n_gpus": 1, "mode": "single", "witness": "the application's gadgets",
        "lagrange_commitments": True, "lagrange_all": True, "feature_set": "shuffle" if shuffle else "default",
        "proof_bytes": len(proof.to_bytes_be()), "public_inputs": len(cs.public_vars_constraint_indices),
        "prove_ms": dt * 1e3, "proofs_per_s": 1 / dt, "e2e_prove_ms": dt_e2e * 1e3, "e2e_proofs_per_s": 1 / dt_e2e,
        "h2d_bytes_per_step": int(wit_host.nbytes) + n_sel * n * 4, "d2h_bytes_per_step": n_msm * 96 + (20 if shuffle else 16) * 32, "steps": K,
        "launches_per_proof": int(launches), "rounds_ms": {k: v / K for k, v in timings.items()},
        "prover": "uzkge_cuda_plonk_prove (compiled, one C-ABI call per proof)", "python_mirror_prove_ms": mirror_ms,
        "ops_per_proof": ops, "ops_source": "counted by the prover (uzkge_plonk_proof)",
        "build_cs_host_s": build_s, "setup_s": setup_s, "refresh_public_key_ms": refresh_ms,
        "deterministic": proof.to_bytes_be() == proof2.to_bytes_be(),
    }
    pcs.close()
    lagrange.close()
    del wit_host
    wit_pinned.free()
    return res


def zshuffle_proofs(dev, K: int, W: int, cards: int = 52) -> dict:
    """BASELINE.json configs[0]: the zshuffle circuit itself (shuffle/src/build_cs.rs:26-56: `cards` remark gadgets + the permutation
    gadget, 2^14 gates for 52 cards) built by the host mirror (uzkge_b200/shuffle.py), `shuffle` feature set, a joint key loaded
    with refresh_prover_params_public_key."""
    from uzkge_b200 import plonk
    from uzkge_b200 import shuffle as sh
    from uzkge_b200.rng import ChaChaRng

    prng = ChaChaRng.from_seed(bytes(32))
    t0 = time.perf_counter()
    apk = sh.rand_point(prng)
    deck = [sh.Ciphertext.rand(prng) for _ in range(cards)]
    cs, _ = sh.build_cs(plonk.TurboCS(), prng, apk, deck)
    return _app_circuit_proofs(dev, K, W, f"zshuffle-{cards}", cs, time.perf_counter() - t0, True, b"Plonk shuffle Proof", cards, apk)


def zmatchmaking_proofs(dev, K: int, W: int) -> dict:
    """BASELINE.json's zmatchmaking circuit (matchmaking/src/build_cs.rs:26-67: N = 50 inputs, 18 Anemoi permutations, 2^13 gates)
    built by uzkge_b200/matchmaking.py, default feature set (the crate's own Cargo.toml does not enable `shuffle`)."""
    from uzkge_b200 import matchmaking as mm
    from uzkge_b200 import plonk
    from uzkge_b200.rng import ChaChaRng, fr_rand

    prng = ChaChaRng.from_seed(bytes(32))
    t0 = time.perf_counter()
    cs, _ = mm.build_cs(plonk.TurboCS(), list(range(1, mm.N + 1)), fr_rand(prng), fr_rand(prng))
    return _app_circuit_proofs(dev, K, W, "zmatchmaking", cs, time.perf_counter() - t0, False, mm.PLONK_PROOF_TRANSCRIPT, mm.N)


# ------------------------------------------------------------------------------------------------ our arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "both", "msm", "ntt", "plonk"])
    ap.add_argument("--plonk-logs", default="14,22", help="log2 circuit sizes of the synthetic TurboPlonK proofs")
    ap.add_argument("--no-zshuffle", action="store_true", help="skip the application circuits (zmatchmaking, zshuffle-52) of the PlonK block")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    from uzkge_b200 import ffi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- uzkge_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
        dist.init_process_group("nccl", device_id=dev)
    ffi.init(local_rank)
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    K, W = args.steps, args.warmup
    launches0 = ffi.launch_count()
    results = {}
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_region0 = time.time()
    if sampler:
        sampler.start()

    # -------------------------------------------------------------------------------------------- MSM
    if args.workload in ("all", "both", "msm"):
        n = 1 << LOG_MSM
        # every rank builds its own slice of the synthetic SRS on its own GPU (setup path, untimed): 2^20 powers-of-tau
        # points with a rank-specific trapdoor, so the N slices are N * 2^20 distinct bases
        tau = random_fr(1, 0xB2000001 + rank)[0]
        t0 = time.time()
        bases = ffi.srs_generate(tau, n)
        t_gen = time.time() - t0
        h = ffi.srs_upload(bases, args.window_bits)
        info = ffi.srs_info(h)
        nsets = 8  # 8 x 32 MiB of scalars > 126 MB L2
        host_sets = [random_fr(n, 0xB2000002 + 97 * rank + i) for i in range(nsets)]
        d_sets = [torch.from_numpy(s.view(np.int64)).to(dev) for s in host_sets]
        d_out = torch.zeros(12 * max(world, 1), dtype=torch.int64, device=dev)
        d_mine = torch.zeros(12, dtype=torch.int64, device=dev)
        d_acc = torch.zeros(12, dtype=torch.int64, device=dev)

        def msm_step(i):
            ffi.msm_g1_device(h, d_sets[i % nsets].data_ptr(), n, d_mine.data_ptr(), sptr)
            if world > 1:
                dist.all_gather_into_tensor(d_out, d_mine)

        def combine():
            # N - 1 projective adds of the gathered partial sums, on the device (one 1-thread kernel), then ONE 96-byte read
            ffi.g1_sum_device(d_out.data_ptr(), world, d_acc.data_ptr(), sptr)
            return d_acc.cpu().numpy().view(np.uint64)

        for i in range(W):
            msm_step(i)
        barrier()
        ffi.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(K):
            msm_step(i)
        e1.record(stream)
        barrier()
        single_ms = max_over_ranks(e0.elapsed_time(e1) / K)      # one API call per MSM: the latency of a single commitment
        prof = ffi.profile_read("msm")
        ffi.profile_enable(False)

        # throughput: the K steps travel in ONE batch call (uzkge_cuda_msm_g1_batch_device): the engine pipelines them over two
        # workspaces -- counting sort of step i + 1 and bucket reduction of step i - 1 under the accumulate kernel of step i
        d_outs = torch.zeros(12 * K, dtype=torch.int64, device=dev)
        d_gath = torch.zeros(12 * K * max(world, 1), dtype=torch.int64, device=dev)
        ptrs = [d_sets[i % nsets].data_ptr() for i in range(K)]

        def msm_batch():
            ffi.msm_g1_batch_device(h, ptrs, [n] * K, d_outs.data_ptr(), sptr)
            if world > 1:
                dist.all_gather_into_tensor(d_gath, d_outs)

        msm_batch()
        barrier()
        e0.record(stream)
        msm_batch()
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / K)
        # the pipelined results must equal the one-by-one results
        chk = torch.zeros(12, dtype=torch.int64, device=dev)
        for i in (0, K - 1):
            ffi.msm_g1_device(h, ptrs[i], n, chk.data_ptr(), sptr)
            a, b = chk.cpu().numpy().view(np.uint64), d_outs[12 * i: 12 * i + 12].cpu().numpy().view(np.uint64)
            if not np.array_equal(ffi.g1_to_affine(a), ffi.g1_to_affine(b)):
                raise SystemExit("bench.py: pipelined batch MSM differs from the single-call MSM -- refusing to report a number")
        if world > 1:
            t0 = time.time()
            combine()
            combine_ms = (time.time() - t0) * 1e3
        else:
            combine_ms = 0.0

        # e2e through the host-pointer ABI: pinned scalars, H2D + MSM + D2H of the 96-byte results inside the timed region.
        # uzkge_cuda_msm_g1_batch takes the K host vectors in one call (copies of later steps run under the kernels of earlier
        # ones); the one-call-per-MSM figure is reported next to it.
        pinned = [ffi.PinnedArray((n, 4)) for _ in range(4)]
        for k, pa in enumerate(pinned):
            pa.array[:] = host_sets[k]
        for i in range(2):
            ffi.msm_g1(h, pinned[i % 4].array)
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            out = ffi.msm_g1(h, pinned[i % 4].array)
            if world > 1:
                d_mine.copy_(torch.from_numpy(out.view(np.int64)))
                dist.all_gather_into_tensor(d_out, d_mine)
                combine()
        barrier()
        e2e_single_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / K)
        vecs = [pinned[i % 4].array for i in range(K)]
        ffi.msm_g1_batch(h, vecs)
        barrier()
        t0 = time.perf_counter()
        outs = ffi.msm_g1_batch(h, vecs)
        if world > 1:
            d_outs.copy_(torch.from_numpy(outs.view(np.int64).reshape(-1)))
            dist.all_gather_into_tensor(d_gath, d_outs)
            ffi.g1_sum_device(d_gath.data_ptr(), world, d_outs.data_ptr(), sptr, k=K)     # K results x (N - 1) additions in one launch
            d_outs.cpu()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / K)
        for pa in pinned:
            pa.free()

        acc_ms = prof["ms"]["accumulate"]
        W_ = info["windows"]
        c_ = info["window_bits"]
        fq_mul = 10.0 * n * W_ + 28.0 * (1 << (c_ - 1))
        acc_fq_mul = 10.0 * max(0, n * W_ - (1 << (c_ - 1)))
        results["msm"] = {
            "ms": ms, "e2e_ms": e2e_ms, "single_ms": single_ms, "e2e_single_ms": e2e_single_ms, "n": n, "phases_ms": prof["ms"],
            "combine_ms": combine_ms,
            "window_bits": c_, "windows": W_, "table_bytes": info["device_bytes"], "precompute_ms": info["precompute_ms"],
            "srs_generate_s": t_gen, "fq_mul": fq_mul, "acc_fq_mul": acc_fq_mul, "acc_ms": acc_ms,
            "host0": host_sets[0], "bases": bases, "handle": h,
        }

    # -------------------------------------------------------------------------------------------- NTT
    if args.workload in ("all", "both", "ntt"):
        n = 1 << LOG_NTT
        nbuf = 4  # 4 x 128 MiB inputs > L2
        hx = random_fr(n, 0xB2000003 + rank)
        d_in = [torch.from_numpy(np.roll(hx, 17 * i, axis=0).view(np.int64)).to(dev) for i in range(nbuf)]
        d_outb = torch.empty(4 * n, dtype=torch.int64, device=dev)
        d_scr = torch.empty(4 * n, dtype=torch.int64, device=dev)

        def ntt_step(i):
            ffi.ntt_fr_device(d_in[i % nbuf].data_ptr(), d_outb.data_ptr(), d_scr.data_ptr(), n, n, False, None, sptr)

        for i in range(W):
            ntt_step(i)
        barrier()
        ffi.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(K):
            ntt_step(i)
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / K)
        prof = ffi.profile_read("ntt")
        ffi.profile_enable(False)
        pinned = ffi.PinnedArray((n, 4))
        pinned.array[:] = hx
        for i in range(2):
            ffi.ntt_fr_inplace(pinned.array, n, n, bool(i & 1))
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            ffi.ntt_fr_inplace(pinned.array, n, n, bool(i & 1))  # forward / inverse alternate: the data stays bounded
        barrier()
        e2e_single_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / K)
        roundtrip_ok = bool(np.array_equal(pinned.array, hx)) if K % 2 == 0 else None
        # the same through uzkge_cuda_ntt_fr_batch: 8 independent host vectors per call, H2D / transform / D2H pipelined
        nb = 8
        pins = [pinned] + [ffi.PinnedArray((n, 4)) for _ in range(nb - 1)]
        for i, pa in enumerate(pins):
            pa.array[:] = np.roll(hx, 5 * i, axis=0)
        bufs = [pa.array for pa in pins]
        ffi.ntt_fr_batch_inplace(bufs, [n] * nb, n)
        ffi.ntt_fr_batch_inplace(bufs, [n] * nb, n, inverse=True)
        roundtrip_ok = roundtrip_ok and bool(np.array_equal(pins[3].array, np.roll(hx, 15, axis=0)))
        calls = max(1, (K + nb - 1) // nb)
        barrier()
        t0 = time.perf_counter()
        for c in range(calls):
            ffi.ntt_fr_batch_inplace(bufs, [n] * nb, n, inverse=bool(c & 1))
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / (calls * nb))
        for pa in pins:
            pa.free()
        results["ntt"] = {"ms": ms, "e2e_ms": e2e_ms, "e2e_single_ms": e2e_single_ms, "n": n, "phases_ms": prof["ms"],
                          "roundtrip_ok": roundtrip_ok, "hx": hx}

    # -------------------------------------------------------------------------------------------- strong scaling: ONE 2^20 MSM over N GPUs
    if args.workload in ("all", "both", "msm") and world > 1:
        n = 1 << LOG_MSM
        lo, hi = rank * n // world, (rank + 1) * n // world
        tau_s = random_fr(1, 0xB2000010)[0]                    # the same trapdoor on every rank: one SRS, rank r uploads its slice
        part = ffi.srs_generate(tau_s, n)[lo:hi]
        hs = ffi.srs_upload(np.ascontiguousarray(part), args.window_bits)
        del part
        sc_full = random_fr(n, 0xB2000011)                     # the same scalars on every rank
        d_sc = torch.from_numpy(sc_full[lo:hi].view(np.int64)).to(dev)
        d_part, d_all = torch.zeros(12, dtype=torch.int64, device=dev), torch.zeros(12 * world, dtype=torch.int64, device=dev)

        def strong_step():
            ffi.msm_g1_device(hs, d_sc.data_ptr(), hi - lo, d_part.data_ptr(), sptr)
            dist.all_gather_into_tensor(d_all, d_part)
            ffi.g1_sum_device(d_all.data_ptr(), world, d_part.data_ptr(), sptr)      # the N - 1 projective additions, on the device

        for _ in range(W):
            strong_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            strong_step()
        e1.record(stream)
        barrier()
        strong_ms = max_over_ranks(e0.elapsed_time(e1) / K)
        # parity: sum_i s_i tau^i G by the trapdoor (rank 0 checks, the verdict is shared)
        ok = True
        if rank == 0:
            from oracle import cpu as oc  # checker only, outside the timed region

            want = oc.g1_to_affine(oc.g1_mul(ffi.srs_generate(tau_s, 1)[0], oc.fr_eval(sc_full, tau_s)))
            ok = bool(np.array_equal(oc.g1_to_affine(d_part.cpu().numpy().view(np.uint64)), want))
        one_ms = results["msm"]["single_ms"]
        results["msm_strong"] = {
            "metric": "bn254_g1_msm_2^20_points_per_s (ONE MSM over N GPUs)", "n_gpus": world, "log_n": LOG_MSM,
            "value": n / (strong_ms * 1e-3), "unit": "points/s", "ms_per_step": strong_ms, "single_gpu_ms": one_ms,
            "speedup": one_ms / strong_ms, "efficiency": one_ms / strong_ms / world, "exchanged_bytes_per_step": 96 * world * world,
            "window_bits": ffi.srs_info(hs)["window_bits"], "parity_ok": ok,
            "what": "strong scaling: ONE 2^20-point MSM, rank r holds SRS points and scalars [r n / N, (r + 1) n / N); per step one MSM "
                    "call per rank, one 96-byte all-gather (NCCL) and N - 1 projective additions on the device; single_gpu_ms = one "
                    "2^20 MSM per call on one GPU of this run (detail.single_call_ms); result checked against the trapdoor",
        }
        ffi.srs_free(hs)
        del d_sc
        torch.cuda.empty_cache()

    # -------------------------------------------------------------------------------------------- distributed NTT (N = 2, 4, 8)
    if args.workload in ("all", "both", "ntt") and world in (2, 4, 8):
        from uzkge_b200 import dist as udist

        reps = max(3, min(K, 10))
        strong = {}
        for lg in (22, 24):
            nt = 1 << lg
            Lr = nt // world
            mine = torch.from_numpy(random_fr(Lr, 0xB2000005 + rank).view(np.int64).reshape(-1)).to(dev)
            # the same size on ONE GPU of this run (rank 0's figure is shared): the numerator of the strong-scaling efficiency
            xin = torch.from_numpy(random_fr(nt, 0xB2000006).view(np.int64).reshape(-1)).to(dev)
            xout, xscr = torch.empty_like(xin), torch.empty_like(xin)
            for _ in range(3):
                ffi.ntt_fr_device(xin.data_ptr(), xout.data_ptr(), xscr.data_ptr(), nt, nt, False, None, sptr)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                ffi.ntt_fr_device(xin.data_ptr(), xout.data_ptr(), xscr.data_ptr(), nt, nt, False, None, sptr)
            e1.record(stream)
            barrier()
            one_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
            del xin, xout, xscr
            y = udist.ntt_fr_distributed(mine, nt, rank, world)
            back = udist.ntt_fr_distributed(y, nt, rank, world, inverse=True)
            rt_ok = bool(torch.equal(back, mine))
            out_dn = {}
            for name, natural in (("natural", True), ("cyclic", False)):
                for _ in range(2):
                    udist.ntt_fr_distributed(mine, nt, rank, world, natural_output=natural)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(reps):
                    udist.ntt_fr_distributed(mine, nt, rank, world, natural_output=natural)
                e1.record(stream)
                barrier()
                out_dn[name] = max_over_ranks(e0.elapsed_time(e1) / reps)
            # both exchanges fused into the cross-rank kernel over peer memory (cudaIpc mappings, NVLink P2P loads / stores)
            peer_err = None
            try:
                peer = udist.PeerNtt(nt, rank, world, dev)
            except Exception as e:  # no peer mappings on this box: the NCCL figures stand (every rank must agree before going on)
                peer, peer_err = None, f"{type(e).__name__}: {e}"
            ok_all = torch.tensor([0.0 if peer is None else 1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
            if ok_all.item() == 1.0:
                peer.x_view.copy_(mine)
                yp = peer.transform()
                rt_ok = rt_ok and bool(torch.equal(yp, udist.ntt_fr_distributed(mine, nt, rank, world, natural_output=False)))
                for _ in range(2):
                    peer.transform()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(reps):
                    peer.transform()
                e1.record(stream)
                barrier()
                out_dn["peer"] = max_over_ranks(e0.elapsed_time(e1) / reps)
                if hasattr(peer, "transform_natural"):
                    yn = peer.transform_natural()
                    rt_ok = rt_ok and bool(torch.equal(yn, y))
                    for _ in range(2):
                        peer.transform_natural()
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(reps):
                        peer.transform_natural()
                    e1.record(stream)
                    barrier()
                    out_dn["peer_natural"] = max_over_ranks(e0.elapsed_time(e1) / reps)
                peer.close()
            else:
                out_dn["peer"] = out_dn["cyclic"]
                peer_err = peer_err or "peer mapping failed on another rank"
            flag = torch.tensor([1.0 if rt_ok else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            best = min(out_dn.values())
            strong[lg] = {
                "metric": f"bn254_fr_ntt_2^{lg}_four_step_elements_per_s", "log_n": lg, "n_gpus": world,
                "value": nt / (out_dn["peer"] * 1e-3), "unit": "elements/s", "ms_per_step": out_dn["peer"],
                "peer_natural_output_ms": out_dn.get("peer_natural"),
                "nccl_natural_output_ms": out_dn["natural"], "nccl_cyclic_output_ms": out_dn["cyclic"], "parity_ok": bool(flag.item() == 1.0),
                "steps": reps, "single_gpu_ms": one_ms, "speedup": one_ms / out_dn["peer"], "efficiency": one_ms / out_dn["peer"] / world,
                "natural_output_efficiency": one_ms / min(out_dn["natural"], out_dn.get("peer_natural", 1e9)) / world,
                "exchanged_bytes_per_step": 2 * 32 * nt * (world - 1) // world, "best_ms": best,
                "what": f"ONE 2^{lg} transform over the N GPUs (four-step).  value / ms_per_step: both exchanges fused into the cross-rank "
                        "kernel over peer memory (dist.PeerNtt: P2P loads from every rank's slice, G-point transforms + twiddles, P2P "
                        "stores into the owners' buffers, local 2^lg/N transform; cyclic output layout).  peer_natural: the local "
                        "transform's last pass stores straight into the owners' natural slices over peer memory (no third exchange).  "
                        "nccl_*: the same steps with NCCL all-to-alls (cyclic = same output layout; natural = one more exchange back to "
                        "contiguous slices); single_gpu_ms: the same size on one GPU in this run; "
                        "parity_ok: inverse(forward) round trip and fused == NCCL, on every rank",
                "peer_memory_error": peer_err,
            }
            del mine, y, back
            torch.cuda.empty_cache()
        results["ntt_distributed"] = strong[24]
        results["ntt_strong"] = strong[22]

    # -------------------------------------------------------------------------------------------- ONE host-pointer transform on all N GPUs
    if args.workload in ("all", "both", "ntt") and world in (2, 4, 8):
        import datetime

        store = dist.distributed_c10d._get_default_store()
        key = "uzkge_group_ntt_done"
        if rank != 0:
            torch.cuda.synchronize()
            store.wait([key], datetime.timedelta(minutes=30))      # this rank's GPU belongs to rank 0's device group meanwhile
        else:
            ffi.configure("virtual_devices", 0)
            if ffi.init_devices(world) != world:
                raise SystemExit("bench.py: the device group does not cover the GPUs of the job")
            n = 1 << LOG_NTT
            hx = results["ntt"]["hx"]
            pin = ffi.PinnedArray((n, 4))
            pin.array[:] = hx
            # the same call on ONE GPU while the other ranks are idle (the per-rank figure above is taken with all N ranks copying at once)
            for i in range(2):
                ffi.ntt_fr_inplace(pin.array, n, n, bool(i & 1))
            t0 = time.perf_counter()
            for i in range(K):
                ffi.ntt_fr_inplace(pin.array, n, n, bool(i & 1))
            one = (time.perf_counter() - t0) * 1e3 / K
            pin.array[:] = hx
            for i in range(4):
                ffi.ntt_fr_multi_inplace(pin.array, n, n, bool(i & 1))
            t0 = time.perf_counter()
            for i in range(K):
                ffi.ntt_fr_multi_inplace(pin.array, n, n, bool(i & 1))     # forward / inverse alternate: the data stays bounded
            group_ms = (time.perf_counter() - t0) * 1e3 / K
            ok = bool(np.array_equal(pin.array, hx)) if K % 2 == 0 else None
            fwd = np.array(pin.array) if K % 2 == 0 else None
            if fwd is not None:
                ffi.ntt_fr_multi_inplace(fwd, n, n, False)
                ok = ok and bool(np.array_equal(fwd, ffi.ntt_fr(hx, n)))        # the group's transform == the single-GPU transform
            pin.free()
            if "msm" in results:
                # the round's independent commitments dealt to the GPUs (plonk/prover.rs:132-192, helpers.rs:1323-1408) and
                # one commitment split by points (row 1), both through the host-pointer calls a Rust caller makes, from ONE process
                rm = results["msm"]
                nm, kk = rm["n"], 8
                vecs = [random_fr(nm, 0xB2000020 + j) for j in range(kk)]
                pins = [ffi.PinnedArray((nm, 4)) for _ in range(kk)]
                for pa, v in zip(pins, vecs):
                    pa.array[:] = v
                arrs = [pa.array for pa in pins]
                h_rep = ffi.srs_upload_multi(rm["bases"], ffi.MULTI_REPLICATED, args.window_bits)
                h_split = ffi.srs_upload_multi(rm["bases"], ffi.MULTI_SPLIT, args.window_bits)

                def wall(fn, reps=5):
                    fn()
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        out = fn()
                    return (time.perf_counter() - t0) * 1e3 / reps, out

                t_one, o_one = wall(lambda: ffi.msm_g1_batch(rm["handle"], arrs))
                t_rep, o_rep = wall(lambda: ffi.msm_g1_batch(h_rep, arrs))
                t_spl, o_spl = wall(lambda: ffi.msm_g1_batch(h_split, arrs))
                t_one1, o_one1 = wall(lambda: ffi.msm_g1(rm["handle"], arrs[0]))
                t_spl1, o_spl1 = wall(lambda: ffi.msm_g1(h_split, arrs[0]))
                same = all(np.array_equal(ffi.g1_to_affine(o_one[j]), ffi.g1_to_affine(o_rep[j])) and
                           np.array_equal(ffi.g1_to_affine(o_one[j]), ffi.g1_to_affine(o_spl[j])) for j in range(kk))
                same = same and bool(np.array_equal(ffi.g1_to_affine(o_one1), ffi.g1_to_affine(o_spl1)))
                results["commit_group_e2e"] = {
                    "n_gpus": world, "points": nm, "commitments_per_round": kk, "parity_ok": bool(same),
                    "one_gpu_round_ms": t_one, "dealt_round_ms": t_rep, "point_split_round_ms": t_spl,
                    "dealt_speedup": t_one / t_rep, "point_split_speedup": t_one / t_spl,
                    "one_gpu_single_commit_ms": t_one1, "point_split_single_commit_ms": t_spl1,
                    "what": f"{kk} independent 2^{LOG_MSM}-point commitments (a prover round) through uzkge_cuda_msm_g1_batch on host scalars, "
                            "from ONE process: on one GPU; dealt to the N GPUs (UZKGE_MULTI_REPLICATED: MSM j on GPU j mod N); every MSM "
                            "split by points over the N GPUs (UZKGE_MULTI_SPLIT, partial sums added on the host).  H2D of the scalars "
                            "(32 MiB per commitment) inside every figure; results compared",
                }
                ffi.srs_free(h_rep)
                ffi.srs_free(h_split)
                for pa in pins:
                    pa.free()
            results["ntt_group_e2e"] = {
                "metric": "bn254_fr_ntt_2^22_elements_per_s (ONE host-pointer transform on N GPUs, one process)", "n_gpus": world,
                "value": n / (group_ms * 1e-3), "unit": "elements/s", "ms_per_step": group_ms, "single_gpu_single_call_ms": one,
                "speedup": one / group_ms, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n, "parity_ok": ok,
                "what": "uzkge_cuda_ntt_fr_multi from ONE process: every GPU uploads / returns 1/N of the vector over its own host link, "
                        "four-step over peer memory in between (cross kernel, local transform storing into the owners' natural "
                        "slices); compared with ONE uzkge_cuda_ntt_fr call on one GPU (copy in, transform, copy out) made while the other "
                        "ranks are idle",
            }
            store.set(key, "1")
        barrier()

    # -------------------------------------------------------------------------------------------- PlonK (rank 0, N = 1)
    if args.workload in ("all", "plonk"):
        results["plonk"] = plonk_proofs(args, dev, K, W, rank, world, barrier, max_over_ranks)
        if world == 1 and not args.no_zshuffle:
            for app in (zmatchmaking_proofs, zshuffle_proofs):
                try:
                    results["plonk"]["sizes"].append(app(dev, K, W))
                except Exception as exc:       # reported in the JSON line, never silently: the synthetic sizes above still stand
                    results["plonk"].setdefault("app_errors", []).append(f"{app.__name__}: {exc!r}")

    t_region1 = time.time()
    clocks = sampler.stop(t_region0, t_region1) if sampler else None
    launches = ffi.launch_count() - launches0

    # integer-pipe roof, measured live twice: by the independent probe (IMAD.WIDE issue rate, no library code) and by the library's
    # own dependent Montgomery-multiplication chains; the roofline's peak is the probe's
    probe = imad_probe(local_rank) if rank == 0 else {}
    fq_peak = ffi.bench_field_mul("fq", 2000)
    imad_peak = probe.get("imad_wide_per_s") or fq_peak * IMAD_PER_MUL
    imad_peak_src = ("scripts/cuda/pipe_probe --json (independent of the library: dependent IMAD.WIDE chains on every SM, CUDA events, "
                     "this run)" if probe else "library loop x 136 (pipe_probe binary missing)")

    # -------------------------------------------------------------------------------------------- CPU baseline (rank 0, N = 1)
    cpu_baseline = None
    cpu_in = results.get("plonk", {}).pop("_cpu_inputs", None)      # arrays: never part of the JSON line
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu as oc  # the checker / CPU arm only

        oc.set_num_threads(len(os.sched_getaffinity(0)))
        cores = oc.num_threads()
        if "msm" in results:
            r = results["msm"]
            m = r["n"]
            t0 = time.time()
            want = oc.msm_g1(r["bases"][:m], r["host0"][:m])
            dt = time.time() - t0
            got = ffi.msm_g1(r["handle"], r["host0"][:m])
            if not np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(want)):
                raise SystemExit("bench.py: GPU MSM result differs from the CPU oracle -- refusing to report a number")
            cpu_baseline = {"value": m / dt, "unit": "points/s", "cores": cores, "kind": "port",
                            "sample": f"one full 2^{LOG_MSM}-point MSM (the workload's bases and first scalar set), {dt:.2f} s; "
                                      "result compared with the GPU's (affine) before timing was accepted"}
        if cpu_in is not None:
            # a complete CPU proof of the same statement by the compiled CPU prover (oracle/cpu_prover.py + oracle.c: arkworks'
            # Pippenger and radix-2 transforms restated, the quotient map over all cores, serial Horner / division as in the
            # reference), accepted as a baseline only if its proof equals the GPU's byte for byte
            from oracle import cpu_prover as cpp
            from oracle import plonk_prover as opp

            acs = cpp.ArrayCS(cpu_in["selectors"], cpu_in["wiring"], cpu_in["boolean"], cpu_in["pub_constraint"], cpu_in["pub_witness"])
            cpcs = cpp.CpuKzg(cpu_in["srs"])
            t0 = time.time()
            cP = cpp.indexer(acs, cpcs)
            idx_s = time.time() - t0
            reps, t0 = 0, time.time()
            while reps < 3 or (time.time() - t0 < 5.0 and reps < 20):
                cproof = cpp.prover(opp.ChaCha(bytes(32)), opp.Transcript(b"bench"), cpcs, acs, cP, cpu_in["witness"])
                reps += 1
            dt = (time.time() - t0) / reps
            if opp.proof_to_bytes_be(cproof) != cpu_in["proof_bytes"]:
                raise SystemExit("bench.py: the CPU prover's proof differs from the GPU's -- refusing to report a number")
            results["plonk"]["cpu_baseline"] = {
                "value": 1.0 / dt, "unit": "proofs/s", "cores": cores, "kind": "port",
                "sample": f"{reps} complete proofs of the 2^{cpu_in['log_n']}-gate synthetic circuit (same circuit, witness, SRS, RNG seed and "
                          f"transcript label as the GPU's first PlonK row), {dt * 1e3:.0f} ms each (indexer {idx_s:.1f} s, not timed); "
                          "proof bytes equal to the GPU's",
                "prove_ms": dt * 1e3, "gpu_prove_ms": cpu_in["gpu_prove_ms"]}
        if "ntt" in results:
            r = results["ntt"]
            t0 = time.time()
            want = oc.ntt_fr(r["hx"], r["n"])
            dt = time.time() - t0
            got = ffi.ntt_fr(r["hx"], r["n"])
            if not np.array_equal(got, want):
                raise SystemExit("bench.py: GPU NTT result differs from the CPU oracle -- refusing to report a number")
            r["cpu"] = {"value": r["n"] / dt, "unit": "elements/s", "cores": cores, "kind": "port",
                        "sample": f"one full 2^22 NTT, {dt:.2f} s; output compared with the GPU's"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    def msm_block(r):
        n_total = r["n"] * world
        traffic, traffic_src = ncu_traffic("msm", "msm_accumulate_kernel")
        acc_imad = r["acc_fq_mul"] / 10.0 * IMAD_PER_MADD
        return {
            "metric": "bn254_g1_msm_2^20_points_per_s", "value": n_total / (r["ms"] * 1e-3), "unit": "points/s",
            "ms_per_step": r["ms"],
            "e2e": {"value": n_total / (r["e2e_ms"] * 1e-3), "unit": "points/s", "ms_per_step": r["e2e_ms"],
                    "h2d_bytes_per_step": 32 * r["n"], "d2h_bytes_per_step": 96},
            # the binding roof: the multiplier (fmaheavy) pipe.  N*W - 2^(c-1) mixed additions (the first point of a bucket is a copy)
            # of 8 M + 2 S each; peak = IMAD.WIDE issue rate measured by the independent probe in this run
            "roofline": {"bound": "imad", "kernel": "msm_accumulate_kernel", "achieved": acc_imad / (r["acc_ms"] * 1e-3) / 1e12,
                         "peak": imad_peak / 1e12, "unit": "T IMAD/s", "frac": acc_imad / (r["acc_ms"] * 1e-3) / imad_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": r["acc_ms"], "peak_source": imad_peak_src,
                         "peak_issue_model": probe.get("model_per_s", 0) / 1e12 or None,
                         "peak_library_loop": fq_peak * IMAD_PER_MUL / 1e12,
                         "algorithmic": {"mixed_additions": r["acc_fq_mul"] / 10.0, "imad_per_mixed_addition": IMAD_PER_MADD,
                                         "fq_products": r["acc_fq_mul"], "bytes": 96.0 * r["n"]},
                         "note": "MSM is integer-pipe bound, never HBM bound (SURVEY 8d): the HBM view is in hbm_roofline"},
            "hbm_roofline": {"bound": "hbm", "kernel": "msm_accumulate_kernel", "achieved": 96.0 * r["n"] / (r["acc_ms"] * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": 96.0 * r["n"] / (r["acc_ms"] * 1e-3) / 1e9 / hbm_peak,
                             "traffic": traffic, "peak_source": peak_src},
            # the same kernel in field products against the library's own multiplier loop (kept for continuity with round 1)
            "int_roofline": {"bound": "imad", "kernel": "msm_accumulate_kernel", "achieved": r["acc_fq_mul"] / (r["acc_ms"] * 1e-3) / 1e9,
                             "peak": fq_peak / 1e9, "unit": "G Fq-mul/s", "frac": r["acc_fq_mul"] / (r["acc_ms"] * 1e-3) / fq_peak,
                             "fq_mul": r["acc_fq_mul"],
                             "peak_source": "uzkge_cuda_bench_field_mul (dependent 136-IMAD Montgomery chains, measured in this run)"},
            "phases_ms": r["phases_ms"], "single_call_ms": r["single_ms"], "single_call_e2e_ms": r["e2e_single_ms"],
            "window_bits": r["window_bits"], "windows": r["windows"],
            "srs_device_bytes": r["table_bytes"], "srs_precompute_ms": r["precompute_ms"], "combine_ms": r["combine_ms"],
        }

    def ntt_block(r):
        lg = LOG_NTT
        passes = sum(1 for k in ("pass0", "pass1", "pass2") if r["phases_ms"][k] > 0)
        top = max(("pass0", "pass1", "pass2"), key=lambda k: r["phases_ms"][k])
        top_ms = r["phases_ms"][top]
        traffic, traffic_src = ncu_traffic("ntt", "ntt_pass_kernel")
        # Fr products the passes execute (ntt.cu): a pass over a 2^b digit runs b radix-2 stages; the last has no twiddle, and in
        # stage h the butterflies with twiddle index 0 (1 of every h) skip the product -- (N/2)(b - 2 + 2^(1-b)) per pass -- plus
        # N inter-pass twiddle products per pass boundary (digits: ntt_plan.h splits log2 N evenly, larger digits first)
        digits = [lg // passes + (1 if i < lg % passes else 0) for i in range(passes)]
        fr_mul = sum(r["n"] / 2 * (b_ - 2 + 2.0 ** (1 - b_)) for b_ in digits) + r["n"] * (passes - 1)
        b = {
            "metric": "bn254_fr_ntt_2^22_elements_per_s", "value": r["n"] * world / (r["ms"] * 1e-3), "unit": "elements/s",
            "ms_per_step": r["ms"],
            "e2e": {"value": r["n"] * world / (r["e2e_ms"] * 1e-3), "unit": "elements/s", "ms_per_step": r["e2e_ms"],
                    "h2d_bytes_per_step": 32 * r["n"], "d2h_bytes_per_step": 32 * r["n"]},
            # north star: HBM GB/s per pass -- and the multiplier pipe, which is what binds a 256-bit transform
            "roofline": {"bound": "hbm", "kernel": f"ntt_pass_kernel ({top})", "achieved": 64.0 * r["n"] / (top_ms * 1e-3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": 64.0 * r["n"] / (top_ms * 1e-3) / 1e9 / hbm_peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": top_ms,
                         "peak_source": peak_src, "passes": passes},
            "int_roofline": {"bound": "imad", "achieved": fr_mul * IMAD_PER_MUL / (r["ms"] * 1e-3) / 1e12,
                             "peak": imad_peak / 1e12, "unit": "T IMAD/s", "frac": fr_mul * IMAD_PER_MUL / (r["ms"] * 1e-3) / imad_peak,
                             "fr_mul": fr_mul, "peak_source": imad_peak_src, "peak_library_loop": fq_peak * IMAD_PER_MUL / 1e12},
            "phases_ms": r["phases_ms"], "e2e_roundtrip_ok": r["roundtrip_ok"], "single_call_e2e_ms": r["e2e_single_ms"],
            "e2e_note": "e2e: uzkge_cuda_ntt_fr_batch, 8 host vectors per call, H2D / transform / D2H on three streams (PCIe full duplex); "
                        "single_call_e2e_ms: one uzkge_cuda_ntt_fr call per transform (copy in, transform, copy out in sequence)",
        }
        if "cpu" in r:
            b["cpu_baseline"] = r["cpu"]
        return b

    if "msm" not in results and "ntt" not in results:
        pl = results["plonk"]
        first = pl["sizes"][-1]
        line = {
            "metric": "turboplonk_synthetic_proofs_per_s", "value": first["proofs_per_s"], "unit": "proofs/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": first["prove_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (256-bit Montgomery, IMAD carry chains)", "data": "synthetic",
            "config": {"workload": f"{first.get('circuit', 'synthetic TurboPlonK circuit')}, 2^{first['log_n']} gates, full prove, witness resident in HBM"},
            "e2e": {"value": first["e2e_proofs_per_s"], "unit": "proofs/s", "h2d_bytes_per_step": first["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": first["d2h_bytes_per_step"]},
            "gpu_launches": int(launches), "clocks": clocks, "plonk": pl,
        }
        emit(line)
        return 0

    head = "msm" if "msm" in results else "ntt"
    blk = msm_block(results["msm"]) if head == "msm" else ntt_block(results["ntt"])
    line = {
        "metric": blk["metric"], "value": blk["value"], "unit": blk["unit"], "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (256-bit Montgomery, IMAD carry chains)", "data": "synthetic",
        "config": headline_config(head, world),
        "e2e": blk["e2e"], "roofline": blk["roofline"], "int_roofline": blk["int_roofline"], "hbm_roofline": blk.get("hbm_roofline"),
        "gpu_launches": int(launches), "clocks": clocks,
        "cpu_baseline": cpu_baseline if head == "msm" else blk.get("cpu_baseline"),
        "detail": {k: v for k, v in blk.items() if k in ("phases_ms", "single_call_ms", "single_call_e2e_ms", "window_bits", "windows", "srs_device_bytes",
                                                   "srs_precompute_ms", "combine_ms")},
    }
    if head == "msm" and "ntt" in results:
        line["ntt"] = ntt_block(results["ntt"])
    if "plonk" in results:
        line["plonk"] = results["plonk"]
    for key in ("msm_strong", "ntt_strong", "ntt_distributed", "ntt_group_e2e", "commit_group_e2e"):
        if key in results:
            line[key] = results[key]
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON: whatever a library writes to file descriptor 1 on the way (NCCL prints its version
    # banner there) is sent to stderr instead
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.exit(main())
