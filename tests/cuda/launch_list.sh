#!/bin/bash
# ncu launch list (device time per launch) of one bench step; -> gpurun_out/<tag>_launches.csv
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r01}
L=${2:-282}   # kernel launches per forward (hm_forward_launch_count); the window holds two forward boundaries, the summary keeps one complete forward
MINE='regex:gemm_bf16|layernorm_kernel|im2col|vit_attention|skinny_linear|head_|mano_|crop_warp'
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graphs"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s $((3 * L)) -c $((2 * L + 40)) --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
