#!/bin/bash
# Store-path probes for the MANO vertex array (one process per variant: a rejected instruction kills the context)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/${1:-r3b}_store_probe.log
: > $L
P=build/bulk_store_probe
for args in "0 96 16 1 1" "1 96 16 1 1" "1 96 16 2 1" "1 96 8 2 1" "1 96 8 2 2" "0 192 16 1 1" "1 192 16 2 1" "1 192 8 2 1" "1 384 8 2 1" "1 384 4 2 1"; do
  timeout 60 $P $args >> $L 2>&1 || echo "  -> rc $? for $args" >> $L
done
P=build/tma_store_probe
for args in "1 96 16 0 4 4" "2 96 16 0 4 4" "2 96 16 2 4 4" "3 96 16 2 4 4" "3 96 32 2 16 4" "3 192 16 2 4 4" "3 384 16 2 4 4"; do
  timeout 60 $P $args >> $L 2>&1 || echo "  -> rc $? for $args" >> $L
done
cat $L
