#!/bin/bash
# usage (on the GPU box, from the repo root): tools/run_bench_n.sh N TAG [extra bench args] — one bench.py run at N GPUs, JSON line to gpurun_out/bench_TAG.json
N=$1; TAG=$2; shift 2
export ZKB_BENCH_VERBOSE=1 ZKB_BENCH_WATCHDOG=${ZKB_BENCH_WATCHDOG:-800}
if [ "$N" = "1" ]; then
  python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
fi
echo rc=$?
grep -v "^W\|^\[W\|^\*\*\*\*\|OMP_NUM" gpurun_out/bench_$TAG.err | tail -40
python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
except Exception as e:
    print("no json line", e); raise SystemExit
drop=("workload","note","sharding","exchange","peak_source","traffic_source","sample","timer","lowered","config","clocks","e2e_note","parity")
def short(o):
    if isinstance(o,dict): return {k:short(v) for k,v in o.items() if k not in drop}
    return o
for k in ("value","ms_per_step","e2e","e2e_pageable","no_table_ms_per_step","witness_like","parity","ntt","sharded_ntt","sharded_quotient","wrapper_replay","msm_split","voter_replay","st_replay","single_process","sweep","bench_wall_s"):
    print(k, json.dumps(short(l.get(k)) if k!="parity" else l.get(k)))
print("roofline", json.dumps({k:v for k,v in l["roofline"].items() if k in ("frac","ms_per_launch","other_ms","share_of_step")}))
PY
