set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest9.log 2>&1; tail -2 gpurun_out/pytest9.log
rm -f gpurun_out/sweep_replay.jsonl
timeout 900 python tests/tools/sweep.py --sections replay --out gpurun_out/sweep_replay.jsonl > gpurun_out/sweep_replay.log 2>&1; tail -3 gpurun_out/sweep_replay.log; cat gpurun_out/sweep_replay.jsonl
