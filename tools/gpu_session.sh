cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest rc=$?"; grep -E "identity|passed|failed|rror" gpurun_out/pytest_gpu20.log | head -20
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_modules.py -m gpu -x -q > gpurun_out/pytest_gpu20b.log 2>&1; echo "pytest(2) rc=$?"; tail -2 gpurun_out/pytest_gpu20b.log
