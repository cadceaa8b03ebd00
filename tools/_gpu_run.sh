cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_binned.py -x -q -m gpu > gpurun_out/pytest_binned.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_binned.log
python tools/bn_timing.py --config cfg2
python tools/bn_timing.py --config x2s1
for cfg in cfg2 x2s1; do
timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_auto_${cfg}.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_auto_${cfg}.json'))
print('$cfg', round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done
