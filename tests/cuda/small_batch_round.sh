#!/bin/bash
# small-batch (video regime) profile: per-stage times at 1-8 crops and an ncu launch list of one eager forward at 2 and 5 crops
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
tag=${1:-sb}
timeout 300 python tests/cuda/small_batch_profile.py 1 2 5 8 > gpurun_out/${tag}_small_batch.log 2>&1; cat gpurun_out/${tag}_small_batch.log | tail -6
MINE='regex:gemm_bf16|layernorm_kernel|im2col|vit_attention|skinny_linear|head_|mano_|crop_warp'
for B in 2 5; do
CMD="python bench.py --batch $B --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graphs"
$CMD > gpurun_out/${tag}_plain_b$B.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 2000 --csv --log-file gpurun_out/${tag}_b${B}_launches.csv $CMD > gpurun_out/${tag}_ncu_b$B.log 2>&1
echo "ncu B=$B rc=$?"
done
