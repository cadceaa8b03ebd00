cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_parity.py -x -q -m gpu -k "metrics or cli" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_m.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
