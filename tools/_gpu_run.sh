cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_binned.py -x -q -m gpu > gpurun_out/pytest_binned.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_binned.log
