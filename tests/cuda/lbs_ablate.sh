cd /root/repo; mkdir -p gpurun_out
for d in 0 2; do echo "dbg=$d"; HM_MANO_DBG=$d timeout 120 python tests/cuda/lbs_profile.py 65536 2>&1 | tail -1; done
