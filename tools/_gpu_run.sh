cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_binned.py -x -q -m gpu -k "matches_oracle or fuzz or randomised" > gpurun_out/pytest_binned.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_binned.log
for cfg in cfg2 x2s1; do
timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t.json 2> gpurun_out/bench_auto.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_t.json'))
print('$cfg', round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done
