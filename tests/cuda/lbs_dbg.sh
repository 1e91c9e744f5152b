#!/bin/bash
# fused MANO kernel: ablations (HM_MANO_DBG bits: 1 no vertex stores, 2 no skinning loop, 4 accumulator read only) and phase clocks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/${1:-lbsdbg}_lbs_dbg.log
: > $L
for d in 0 1 2 3 7; do
  echo "== HM_MANO_DBG=$d" >> $L
  HM_MANO_DBG=$d HM_MANO_CLOCKS=1 timeout 120 python tests/cuda/lbs_profile.py 65536 2>&1 | tail -6 >> $L
done
cat $L
