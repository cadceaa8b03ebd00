cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ch in 1 2; do
MULUT_HOST_CHUNK=$ch timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_chunk$ch.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_chunk$ch.json'))
print('chunk $ch', round(d['value']), round(d['e2e']['value']))
PY
done
