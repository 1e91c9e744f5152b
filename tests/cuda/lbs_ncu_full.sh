# usage (GPU box): bash tests/cuda/lbs_ncu_full.sh <tag>  -> one ncu --set full capture of the fused MANO kernel (+ prep)
tag=${1:-lbs}
cd /root/repo
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"mano_fused|mano_prep" -s 4 -c 2 -f -o gpurun_out/${tag}_lbs_full python tests/cuda/lbs_profile.py 65536 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out/${tag}_lbs_full.ncu-rep
