cd /root/repo; mkdir -p gpurun_out
for k in 1280 5120; do
    echo "== N=1280 K=$k cta_group=2"
    timeout 90 ./build/gemm_check 2 256 2 12288 1280 $k 10 | cut -c1-60,100-230
done
bash tests/cuda/check_round.sh ${1:-r01w}
