cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in cfg2 x2s1 cfg3; do for k in auto quad; do
timeout 300 python bench.py --config $cfg --data natural --kernel $k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nat_${cfg}_$k.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_nat_${cfg}_$k.json'))
print('$cfg $k natural', round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})
PY
done; done
