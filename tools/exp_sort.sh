#!/bin/bash
# A/B of the MSM's own bucket sort (default) against the toolkit's radix sort (ZKB_MSM_SORT=cub): per-phase times by scalar distribution
for mode in own cub; do
  for dist in U W E; do
    echo "--- sort=$mode dist=$dist"
    for k in 20 22 24; do ZKB_MSM_SORT=$mode timeout 600 python tools/profile_run.py msm --log-n $k --reps 3 --dist $dist 2>&1 | tail -1 | cut -c1-330; done
  done
done
