#!/bin/bash
# ncu --set full captures of every kernel family of the path, one build (each only after the same command exited 0 plain):
#   ncu_round.sh <tag>   ->  gpurun_out/<tag>_{gemm,misc,head,mano,crop}.ncu-rep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=${1:-r3n}
NCU="ncu --set full --clock-control none --import-source on"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graphs"
$CMD > gpurun_out/${T}_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
# one ViT block's four GEMMs (qkv, proj, fc1, fc2) of a timed step
$NCU -k regex:gemm_bf16_tn_kernel -s 140 -c 4 -o gpurun_out/${T}_gemm $CMD > gpurun_out/${T}_ncu_gemm.log 2>&1; echo "gemm rc=$?"
$NCU -k regex:"vit_attention|layernorm_kernel" -s 40 -c 3 -o gpurun_out/${T}_misc $CMD > gpurun_out/${T}_ncu_misc.log 2>&1; echo "misc rc=$?"
$NCU -k regex:"head_|skinny_linear" -s 25 -c 12 -o gpurun_out/${T}_head $CMD > gpurun_out/${T}_ncu_head.log 2>&1; echo "head rc=$?"
python tests/cuda/lbs_profile.py 65536 > gpurun_out/${T}_lbs_plain.log 2>&1 &&
$NCU -k regex:mano_fused -s 4 -c 1 -o gpurun_out/${T}_mano python tests/cuda/lbs_profile.py 65536 > gpurun_out/${T}_ncu_mano.log 2>&1; echo "mano rc=$?"
python tests/cuda/crop_profile.py 64 320 > gpurun_out/${T}_crop_plain.log 2>&1 && cat gpurun_out/${T}_crop_plain.log | tail -1 &&
$NCU -k regex:crop_warp -s 4 -c 1 -o gpurun_out/${T}_crop python tests/cuda/crop_profile.py 64 320 > gpurun_out/${T}_ncu_crop.log 2>&1; echo "crop rc=$?"
