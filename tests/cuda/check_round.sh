# usage (on the GPU box): bash tests/cuda/check_round.sh <tag>   -> GPU tests + short bench, logs in gpurun_out/
tag=${1:-chk}
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s -x > gpurun_out/${tag}_pytest.log 2>&1
grep -E "parity:|passed|failed|FAILED|ERROR|Error|assert " gpurun_out/${tag}_pytest.log | head -20
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['per_family_tflops']); print(d['stage_ms_per_step']); print(d['mano_lbs'], d['clocks'])
PY
