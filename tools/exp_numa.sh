#!/bin/bash
# e2e of the weak-scaled headline with / without binding every rank to its GPU's NUMA node (N from $1)
N=${1:-8}
for numa in 1 0; do
  echo "--- ZKB_BENCH_NUMA=$numa"
  ZKB_BENCH_NUMA=$numa python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --skip-ntt --skip-cpu --skip-replay 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('value',l['value'],'ms',l['ms_per_step'],'e2e',l['e2e']['value'],l['e2e'].get('ms_per_step'),'pageable',l['e2e_pageable']['value'],'numa',l['config'].get('host_numa'))"
done
nproc; ls /sys/devices/system/node/ | head; nvidia-smi topo -m 2>/dev/null | head -14
