# usage (GPU box): bash tests/cuda/lbs_round.sh <tag>  -> MANO parity tests + config[3] timing + launch list
tag=${1:-lbs}
cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 60 -p no:cacheprovider -x -k "mano" > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
timeout 120 python tests/cuda/lbs_profile.py 65536 > gpurun_out/${tag}_lbs.log 2>&1; tail -3 gpurun_out/${tag}_lbs.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mano|gemm_bf16" -c 40 --csv --log-file gpurun_out/${tag}_lbs_launches.csv python tests/cuda/lbs_profile.py 65536 > gpurun_out/${tag}_ncu.log 2>&1
python tests/tools/summarise_profiles.py ${tag}_lbs 2>/dev/null | tail -8
