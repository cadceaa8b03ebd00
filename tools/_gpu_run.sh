cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/finetune_bench.py > gpurun_out/finetune_n2.json 2> gpurun_out/ft2.err; echo rc=$?; tail -2 gpurun_out/ft2.err; grep workload gpurun_out/finetune_n2.json
for rep in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/bench_n2.json') if l.strip()]
print(len(lines), 'stdout lines')
d=json.loads(lines[-1]); print('n=2', round(d['value']), round(d['e2e']['value']), d['n_gpus'], d['gpu_launches'])
PY
done
