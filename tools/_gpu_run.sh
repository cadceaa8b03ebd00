cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_finetune.py -x -q -m gpu > gpurun_out/pytest_ft.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_ft.log
timeout 300 python tools/finetune_bench.py 2>gpurun_out/ft.err | tail -1
timeout 300 python tools/finetune_bench.py --smooth 2>gpurun_out/ft.err | tail -1
