cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_binned.py -x -q -m gpu > gpurun_out/pytest_binned.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_binned.log
for cfg in "cfg2 uniform" "x2s1 uniform" "cfg2 natural" "x2s1 natural"; do set -- $cfg
timeout 300 python bench.py --config $1 --data $2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t.json 2> gpurun_out/bench_auto.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_t.json'))
print('$1 $2', round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done
