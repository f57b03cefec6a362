#!/bin/bash
# round-2 profile captures (one GPU).  Each ncu run only after the same command exited 0 without ncu.
set -x
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-replay > gpurun_out/plain_bench_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-replay > gpurun_out/ncu_bench_r2.log 2>&1
python tools/profile_run.py msm --log-n 24 --reps 1 > gpurun_out/plain_msm24_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_direct_kernel -c 1 -o gpurun_out/prof_msm_acc_r2 python tools/profile_run.py msm --log-n 24 --reps 1 > gpurun_out/ncu_msm_r2.log 2>&1
python tools/profile_run.py ntt --log-n 22 --cols 16 --reps 1 > gpurun_out/plain_ntt22_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 3 -o gpurun_out/prof_ntt_r2 python tools/profile_run.py ntt --log-n 22 --cols 16 --reps 1 > gpurun_out/ncu_ntt_r2.log 2>&1
python tools/profile_run.py msm --log-n 16 --reps 1 > gpurun_out/plain_msm16_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"msm_accumulate_kernel|msm_sum_tree_kernel|msm_reduce_segment_kernel" -s 6 -c 6 -o gpurun_out/prof_msm_small_r2 python tools/profile_run.py msm --log-n 16 --reps 1 > gpurun_out/ncu_msm16_r2.log 2>&1
ls -la gpurun_out/*_r2.ncu-rep
tail -3 gpurun_out/plain_msm24_r2.log gpurun_out/plain_ntt22_r2.log gpurun_out/plain_msm16_r2.log
