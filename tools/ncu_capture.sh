set -x
python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/plain_bench_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_bench_b.log 2>&1
python tools/profile_run.py msm --log-n 24 --reps 1 > gpurun_out/plain_msm24.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -c 1 -o gpurun_out/prof_msm_acc_r1b python tools/profile_run.py msm --log-n 24 --reps 1 > gpurun_out/ncu_msm_b.log 2>&1
python tools/profile_run.py ntt --log-n 22 --cols 16 --reps 1 > gpurun_out/plain_ntt22b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 3 -c 3 -o gpurun_out/prof_ntt_r1b python tools/profile_run.py ntt --log-n 22 --cols 16 --reps 1 > gpurun_out/ncu_ntt_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
