cd /root/repo
mkdir -p gpurun_out
(for args in "2 256 0 12288 5120 1280 10" "2 256 1 12288 5120 1280 10" "2 128 1 12288 5120 1280 10" "1 256 1 12288 5120 1280 10" "2 256 0 12288 3840 1280 10" "2 256 3 8192 2432 320 10" "2 128 3 8192 2432 320 10"; do timeout 120 ./build/gemm_check $args; done) > gpurun_out/r01e_gemm.log 2>&1
cut -c1-60,100-230 gpurun_out/r01e_gemm.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s -x > gpurun_out/r01e_pytest.log 2>&1
grep -E "parity:|passed|failed|FAILED|ERROR|Error|assert " gpurun_out/r01e_pytest.log | head -20
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01e_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['per_family_tflops']); print(d['stage_ms_per_step']); print(d['mano_lbs'], d['clocks'])
PY
