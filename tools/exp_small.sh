#!/bin/bash
for ch in 0 26 35 52 64 104; do echo "--- chunk $ch"; python tools/msm_small_times.py --log-n 19 20 21 --chunk $ch --table-only 2>&1 | grep log_n; done
echo "--- chunk 52, tree at <= 2 waves"; ZKB_MSM_TREE_WAVES=2 python tools/msm_small_times.py --log-n 19 20 --chunk 52 --table-only 2>&1 | grep log_n
