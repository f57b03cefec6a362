set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "graph or widened" > gpurun_out/pytest8.log 2>&1; tail -2 gpurun_out/pytest8.log
for c in 1 4; do timeout 300 python tools/profile_run.py graph --log-n 24 --cols $c --reps 5; done > gpurun_out/graph_times2.log 2>&1
timeout 300 python tools/profile_run.py graph --log-n 22 --cols 16 --reps 5 >> gpurun_out/graph_times2.log 2>&1
cat gpurun_out/graph_times2.log
timeout 900 python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/bench9.log 2> gpurun_out/bench9.err; tail -1 gpurun_out/bench9.log; tail -3 gpurun_out/bench9.err
