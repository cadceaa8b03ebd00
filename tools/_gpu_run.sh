cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
for cfg in cfg2 cfg3; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_auto_${cfg}.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_auto_${cfg}.json'))
print('$cfg', round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done
