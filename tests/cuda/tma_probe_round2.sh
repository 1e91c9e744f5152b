cd /root/repo; L=gpurun_out/r3c_tma_probe.log; : > $L
for args in "3 24 16 2 16 1" "3 24 16 2 16 2" "3 24 16 2 16 4" "3 24 16 2 8 2" "3 32 16 2 16 2" "3 48 16 2 8 2" "3 24 8 2 16 2"; do
  timeout 60 build/tma_store_probe $args >> $L 2>&1 || echo "  -> rc $? for $args" >> $L
done
cat $L
