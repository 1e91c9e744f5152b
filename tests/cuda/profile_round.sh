#!/bin/bash
# GPU box: tests + bench line + ncu launch list (+ optional full ncu captures).
# (recipe: /opt/skills/guides/B200_PROFILING.md; each ncu run only after the same command exited 0)
#   profile_round.sh <tag> [full]
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r01}
FULL=${2:-}
MINE='regex:gemm_bf16|layernorm_kernel|im2col|vit_attention|skinny_linear|head_|mano_|crop_warp'
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${TAG}_smi.csv 2>&1
bash tests/cuda/run_gemm_checks.sh "2" > /dev/null 2>&1; cp gpurun_out/gemm_check.log gpurun_out/${TAG}_gemm_check.log
grep -E '"test"' gpurun_out/${TAG}_gemm_check.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "parity:|passed|failed|FAILED|ERROR" gpurun_out/${TAG}_pytest.log | head -20
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 3200 gpurun_out/${TAG}_bench.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --no-graphs"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -s 1100 -c 640 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
if [ -n "$FULL" ]; then
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 140 -c 4 \
    -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu gemm rc=$?"
$CMD > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mano_|vit_attention|layernorm_kernel|skinny_linear|head_" -s 40 -c 30 \
    -o gpurun_out/${TAG}_misc $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu misc rc=$?"
fi
