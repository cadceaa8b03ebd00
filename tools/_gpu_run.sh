cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_finetune.py -x -q -m gpu > gpurun_out/pytest_ft.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_ft.log
timeout 300 python tools/finetune_bench.py > gpurun_out/finetune_graph.json 2>gpurun_out/ft.err; cat gpurun_out/finetune_graph.json
timeout 300 python tools/finetune_bench.py --smooth > gpurun_out/finetune_graph_smooth.json 2>gpurun_out/ft.err; cat gpurun_out/finetune_graph_smooth.json
timeout 300 python tools/finetune_bench.py --eager > gpurun_out/finetune_eager.json 2>gpurun_out/ft.err; cat gpurun_out/finetune_eager.json
