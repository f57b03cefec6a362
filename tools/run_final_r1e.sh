set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest12.log 2>&1; tail -3 gpurun_out/pytest12.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1e.log 2> gpurun_out/bench_r1e.err; tail -1 gpurun_out/bench_r1e.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1e_ref.log 2>&1; tail -1 gpurun_out/bench_r1e_ref.log | cut -c1-200
rm -f gpurun_out/sweep_replay_e.jsonl
timeout 900 python tests/tools/sweep.py --sections replay --out gpurun_out/sweep_replay_e.jsonl > gpurun_out/sweep_replay_e.log 2>&1; cat gpurun_out/sweep_replay_e.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
