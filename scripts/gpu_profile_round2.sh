#!/bin/bash
# Round-2 ncu evidence, one GPU (run under gpurun from the repo root): launch lists and --set full captures of the final kernels.
# Each capture follows a plain run of the same command that exited 0.  Outputs under gpurun_out/prof/.
set -u
O=gpurun_out/prof
mkdir -p $O
NCU="ncu --clock-control none"
run() { echo "== $*"; timeout 600 "$@"; echo "rc=$?"; }

# 1. launch list of the bench's MSM + NTT steps
run python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload both > $O/bench_plain.json 2> $O/bench_plain.err \
  && run $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload both > $O/ncu_bench.log 2>&1

# 2. the MSM kernels at 2^20 (accumulate, sort, reduction)
run python scripts/gpu_msm_once.py 20 > $O/msm_plain.log 2>&1 \
  && run $NCU --set full --import-source on -k regex:"msm_accumulate_kernel|msm_count_scatter|msm_strips|msm_sums|msm_leaves" -s 20 -c 14 -f -o $O/r2b_msm python scripts/gpu_msm_once.py 20 > $O/ncu_msm.log 2>&1
ncu -i $O/r2b_msm.ncu-rep --page raw --csv > $O/r2b_msm_raw.csv 2>/dev/null
ncu -i $O/r2b_msm.ncu-rep --page details > $O/r2b_msm_details.txt 2>/dev/null

# 3. the three passes of a 2^22 transform
run python scripts/gpu_ntt_once.py 22 > $O/ntt_plain.log 2>&1 \
  && run $NCU --set full --import-source on -k regex:"ntt_pass_kernel" -s 3 -c 3 -f -o $O/r2b_ntt python scripts/gpu_ntt_once.py 22 > $O/ncu_ntt.log 2>&1
ncu -i $O/r2b_ntt.ncu-rep --page raw --csv > $O/r2b_ntt_raw.csv 2>/dev/null
ncu -i $O/r2b_ntt.ncu-rep --page details > $O/r2b_ntt_details.txt 2>/dev/null

# 4. the prover's own kernels at n = 2^20 (quotient map, permutation terms, scans, Horner)
run python scripts/gpu_plonk_once.py 20 2 > $O/plonk_plain.log 2>&1 \
  && run $NCU --set full -k regex:"plonk_quotient_kernel|plonk_perm_terms|horner|prod_|grand_product|fr_lincomb" -c 24 -f -o $O/r2b_plonk python scripts/gpu_plonk_once.py 20 1 > $O/ncu_plonk.log 2>&1
ncu -i $O/r2b_plonk.ncu-rep --page raw --csv > $O/r2b_plonk_raw.csv 2>/dev/null
ncu -i $O/r2b_plonk.ncu-rep --page details > $O/r2b_plonk_details.txt 2>/dev/null

# 5. one zshuffle-52 proof by the compiled prover: launch list
run python scripts/gpu_zshuffle_profile.py ncu > $O/zshuffle_plain.log 2>&1 \
  && run $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/r2b_zshuffle_launches.csv python scripts/gpu_zshuffle_profile.py ncu > $O/ncu_zshuffle.log 2>&1
# 6. where the host spends a zshuffle-52 proof (cProfile around the compiled prover's caller)
run python scripts/gpu_zshuffle_profile.py host > $O/r2b_zshuffle_host_profile.txt 2>&1
# the reports themselves stay on the box (gpurun_out/ carries at most 64 MiB back): the exported pages above are what is kept
rm -f $O/*.ncu-rep $O/*.ncu-rep.tmp
ls -la $O
