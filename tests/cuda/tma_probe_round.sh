#!/bin/bash
# TMA-store probe variants (one process each: a rejected instruction kills the context)
P=build/tma_store_probe
L=gpurun_out/${1:-r2e}_tma_probe.log
: > $L
for args in "1 96 16 0 4 4" "2 96 16 0 4 4" "2 96 16 2 4 4" "3 96 16 0 4 4" "3 96 32 0 4 4" "3 96 32 0 16 4" "3 96 32 0 4 8" "3 192 16 0 4 4" "3 32 32 0 4 4" "3 96 64 0 4 2" "3 96 128 0 2 2"; do
  timeout 60 $P $args >> $L 2>&1 || echo "  -> rc $? for $args" >> $L
done
cat $L
