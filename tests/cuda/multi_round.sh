#!/bin/bash
# N-GPU box: the 2-GPU tests (hm_allgather through the C ABI) and the weak / strong scaling bench lines
#   multi_round.sh <tag> <N>
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T=${1:-r3m}; N=${2:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_abi_and_host.py -q -p no:cacheprovider --timeout 300 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-extras > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo "weak rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-cpu-baseline --scaling strong > gpurun_out/${T}_bench_${N}gpu_strong.json 2> gpurun_out/${T}_bench_${N}gpu_strong.err; echo "strong rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${T}_bench_${N}gpu.json', 'gpurun_out/${T}_bench_${N}gpu_strong.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['scaling'], round(d['value'], 1), round(d['ms_per_step'], 3), round(d['e2e']['value'], 1), d['config'].get('global_batch'), d.get('collective'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
