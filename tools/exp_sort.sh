#!/bin/bash
# Per-phase times of a commit by scalar distribution.  The A/B against the toolkit's radix sort recorded in profiles/r2_sort_ab.txt was
# taken with this script while msm.cu still carried that path (ZKB_MSM_SORT=cub, commit 29c0480); it has since been removed — no
# library kernel is left in the package — so the script now times the own sort only.
for dist in U W E; do
  echo "--- dist=$dist"
  for k in 20 22 24; do timeout 600 python tools/profile_run.py msm --log-n $k --reps 3 --dist $dist 2>&1 | tail -1 | cut -c1-330; done
done
