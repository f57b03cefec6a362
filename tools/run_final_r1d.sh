# final round-1 evidence after the quotient-evaluation work (one GPU)
set -x
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1d.log 2> gpurun_out/bench_r1d.err; tail -1 gpurun_out/bench_r1d.log | cut -c1-300
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1d_ref.log 2>&1; tail -1 gpurun_out/bench_r1d_ref.log | cut -c1-400
timeout 900 python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/plain_bench_d.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_bench_d.log 2>&1
timeout 300 python tools/profile_run.py graph --log-n 24 --cols 4 --reps 3 > gpurun_out/plain_graph_d.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_evaluate_kernel -s 2 -c 1 -o gpurun_out/prof_graph_r1d python tools/profile_run.py graph --log-n 24 --cols 4 --reps 3 > gpurun_out/ncu_graph_d.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1d.csv
