set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest3.log 2>&1; tail -3 gpurun_out/pytest3.log
for t in 1 0; do ZKB_NTT_PASS_TABLES=$t python tools/profile_run.py ntt --log-n 22 --cols 16; ZKB_NTT_PASS_TABLES=$t python tools/profile_run.py ntt --log-n 15 --cols 256;  ZKB_NTT_PASS_TABLES=$t python tools/profile_run.py c2e --log-n 22 --cols 4; done
for m in 3 4 5 6; do for g in 8 32; do ZKB_MSM_LOG_M=$m ZKB_MSM_SUM_GROUP=$g python tools/profile_run.py msm --log-n 22 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('m',$m,'g',$g,d['ms'],d['msm_reduce'])"; done; done
for m in 4 5 6; do ZKB_MSM_LOG_M=$m ZKB_MSM_SUM_GROUP=8 python tools/profile_run.py msm --log-n 24 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('24 m',$m,d['ms'],d['msm_reduce'])"; done
ZKB_MSM_LOG_M=4 ZKB_MSM_SUM_GROUP=8 python tools/profile_run.py msm --log-n 16 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('16 m4',d['ms'],d['msm_reduce'])"
