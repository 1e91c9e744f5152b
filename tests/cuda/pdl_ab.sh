cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider 2>&1 | tail -3
echo "== PDL on"; timeout 300 python tests/cuda/small_batch_profile.py 1 2 5 8 2>&1 | grep "graph replay" | cut -c1-42
echo "== PDL off"; HM_NO_PDL=1 timeout 300 python tests/cuda/small_batch_profile.py 1 2 5 8 2>&1 | grep "graph replay" | cut -c1-42
echo "== bench PDL on"; python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
echo "== bench PDL off"; HM_NO_PDL=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
