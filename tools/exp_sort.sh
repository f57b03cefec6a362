#!/bin/bash
# A/B of the MSM's own bucket sort (default) against the toolkit's radix sort (ZKB_MSM_SORT=cub): per-phase times, parity by known dlog
run() { timeout 600 python tools/msm_small_times.py --log-n $1 --table-only 2>&1 | grep "log_n\|rror\|ssert" | sed 's/"wall_ms[^,]*, //' | cut -c1-300; }
echo "--- own, 512 threads x 16"; run "13 16 18 20 22 24"
echo "--- own, 1024 threads x 16"; ZKB200_LIB=$PWD/zksnap-circuits-halo2_b200/libzkb200_t1024.so run "16 20 22 24"
echo "--- own 512, b1 = 10"; ZKB_MSM_SORT_B1=10 run "22 24"
