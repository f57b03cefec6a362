# quotient-evaluation kernel: tests, timings, one ncu --set full capture (run under gpurun, one GPU)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest7.log
for c in 1 4; do timeout 300 python tools/profile_run.py graph --log-n 24 --cols $c --reps 5; done > gpurun_out/graph_times.log 2>&1
timeout 300 python tools/profile_run.py graph --log-n 22 --cols 16 --reps 5 >> gpurun_out/graph_times.log 2>&1
timeout 300 python tools/profile_run.py graph --log-n 20 --cols 1 --reps 5 >> gpurun_out/graph_times.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_evaluate_kernel -s 2 -c 1 -o gpurun_out/prof_graph_r1 python tools/profile_run.py graph --log-n 24 --cols 1 --reps 3 > gpurun_out/ncu_graph.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/pytest7.log; cat gpurun_out/graph_times.log
