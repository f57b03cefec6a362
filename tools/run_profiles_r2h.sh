#!/bin/bash
# round-2 (own bucket sort) profile captures, one GPU.  Each ncu run only after the same command exited 0 without ncu.
set -x
tools/run_bench_n.sh 1 r2h_n1 > gpurun_out/bench_r2h_n1.log 2>&1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-replay > gpurun_out/plain_bench_r2h.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2h.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-replay > gpurun_out/ncu_bench_r2h.log 2>&1
python tools/profile_run.py msm --log-n 24 --reps 1 > gpurun_out/plain_msm24_r2h.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"bsort_count_kernel|bsort_scatter_kernel" -c 4 -o gpurun_out/prof_bsort_r2h python tools/profile_run.py msm --log-n 24 --reps 1 > gpurun_out/ncu_bsort_r2h.log 2>&1
tail -2 gpurun_out/plain_msm24_r2h.log; tail -30 gpurun_out/bench_r2h_n1.log | cut -c1-600
