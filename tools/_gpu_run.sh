cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_binned.py -x -q -m gpu > gpurun_out/pytest_binned.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_binned.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print(round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'], d['gather_roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
