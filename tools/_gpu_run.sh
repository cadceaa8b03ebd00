cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
for f in 16 1; do
timeout 300 python bench.py --frames $f --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f$f.json 2> gpurun_out/bench_auto.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_f$f.json'))
print('frames $f', round(d['value']), round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), d['gpu_launches'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
done
timeout 300 python bench.py --config x2s1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x2s1.json 2> gpurun_out/bench_auto.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_x2s1.json'))
print('x2s1', round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})
PY
