cd /root/repo
for sm in 0 768 1536; do echo "HM_SMALL_M=$sm"; HM_SMALL_M=$sm timeout 300 python tests/cuda/small_batch_profile.py 1 2 3 4 6 8 2>&1 | grep "graph replay" | cut -c1-40; done
