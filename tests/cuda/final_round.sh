# usage (GPU box): bash tests/cuda/final_round.sh <tag>  -> what the driver runs (tests, smoke, bench, reference arm) + launch list
tag=${1:-r01z}
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"; tail -c 600 gpurun_out/${tag}_bench_reference.json
timeout 300 python bench_video.py > gpurun_out/${tag}_video.json 2> gpurun_out/${tag}_video.err; echo "video rc=$?"
bash tests/cuda/launch_list.sh ${tag}
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d.get('cpu_baseline'))
PY
