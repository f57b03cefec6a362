set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "slices or msm or commit or thread" > gpurun_out/pytest10.log 2>&1; tail -2 gpurun_out/pytest10.log
echo "=== overlap on"; timeout 600 python tools/msm_slices_ab.py 24 2>&1 | tail -16
echo "=== overlap off"; ZKB_MSM_OVERLAP=0 timeout 600 python tools/msm_slices_ab.py 24 2>&1 | tail -16
echo "=== 2^22 overlap on"; timeout 600 python tools/msm_slices_ab.py 22 2>&1 | tail -16
echo "=== 2^22 overlap off"; ZKB_MSM_OVERLAP=0 timeout 600 python tools/msm_slices_ab.py 22 2>&1 | tail -16
