#!/bin/bash
# Runs the stand-alone tcgen05 GEMM checks; each case in its own process under a timeout so a
# protocol bug cannot hang the box. Output -> gpurun_out/gemm_check.log
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
LOG=gpurun_out/gemm_check.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run() {
  echo "== gemm_check $*" >> $LOG
  timeout 90 ./build/gemm_check "$@" >> $LOG 2>&1
  echo "exit=$?" >> $LOG
}
CASES=${1:-"1 2"}
for cg in $CASES; do
  # small: one tile, a few k-blocks
  run $cg 256 0 128 256 128 0
  run $cg 256 0 256 512 256 0
  # ragged M (B=1 -> 192 rows)
  run $cg 256 0 192 1280 768 0
  run $cg 256 3 192 1280 768 0
  run $cg 128 2 192 1280 1280 0
  # all epilogues, medium
  run $cg 256 1 1536 5120 1280 5
  run $cg 256 2 1536 1280 5120 5
  run $cg 128 2 1536 1280 5120 5
  # full B=64 shapes, timed
  run $cg 256 0 12288 3840 1280 10
  run $cg 256 1 12288 5120 1280 10
  run $cg 256 2 12288 1280 5120 10
  run $cg 128 2 12288 1280 5120 10
  run $cg 256 2 12288 1280 1280 10
  run $cg 128 2 12288 1280 1280 10
  run $cg 256 3 12288 1280 768 10
done
grep -E '"test"|exit=|error|timed out|mismatch' $LOG | tail -80
