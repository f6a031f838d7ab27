cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu32.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu32.log | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench28.json 2> gpurun_out/bench28.err; echo "bench rc=$?"; cut -c1-180 gpurun_out/bench28.json
