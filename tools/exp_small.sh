#!/bin/bash
python tools/msm_small_times.py --log-n 13 14 15 16 17 18 19 20 22 24 2>&1 | grep "table\": true"
echo "--- tree always (ZKB_MSM_TREE_WAVES=1000)"
ZKB_MSM_TREE_WAVES=1000 python tools/msm_small_times.py --log-n 22 24 2>&1 | grep "table\": true"
echo "--- direct from 2 waves"
ZKB_MSM_TREE_WAVES=1 python tools/msm_small_times.py --log-n 19 20 22 2>&1 | grep "table\": true"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "msm or commit or srs or batch" 2>&1 | tail -3
