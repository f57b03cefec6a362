#!/bin/bash
# round-2 (own bucket sort) profile captures, one GPU.  Each ncu run only after the same command exited 0 without ncu.
set -x
tools/run_bench_n.sh 1 r2m_n1 > gpurun_out/bench_r2m_n1.log 2>&1
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-replay > gpurun_out/plain_bench_r2m.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2m.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-replay > gpurun_out/ncu_bench_r2m.log 2>&1


tail -30 gpurun_out/bench_r2m_n1.log | cut -c1-600
