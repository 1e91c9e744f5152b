#!/bin/bash
# A/B of MANO fused-kernel builds / switches: parity against the oracle, then BASELINE configs[3] timing
#   lbs_ab.sh <tag> <variant> [<variant> ...]    variant = <lib>[,ENV=VAL...]; lib "default" = the in-tree libhamer_b200.so,
#   anything else a path to an A/B build (HAMER_B200_LIB override), e.g.  default  default,HM_MANO_STG=1  build/lib_x.so
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
tag=${1:-lbsab}; shift
L=gpurun_out/${tag}_lbs_ab.log
: > $L
for variant in "$@"; do
  IFS=',' read -ra parts <<< "$variant"
  envs=()
  if [ "${parts[0]}" != default ]; then envs+=("HAMER_B200_LIB=$PWD/${parts[0]}"); fi
  for kv in "${parts[@]:1}"; do envs+=("$kv"); done
  echo "== $variant" >> $L
  env "${envs[@]}" timeout 200 python -u tests/cuda/lbs_check.py 1 2 37 300 1001 2>&1 | grep -E "verts|tail|done|Error|error" >> $L
  for i in 1 2; do env "${envs[@]}" timeout 120 python tests/cuda/lbs_profile.py 65536 2>&1 | tail -1 >> $L; done
done
cat $L
