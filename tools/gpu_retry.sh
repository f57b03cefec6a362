#!/bin/bash
# usage: tools/gpu_retry.sh [gpurun flags ...] -- 'command'   — retries while the pod answers busy (exit code 3 / transient)
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy" ; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
