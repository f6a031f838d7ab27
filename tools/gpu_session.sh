cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu22.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu22.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo "bench rc=$?"; tail -3 gpurun_out/bench19.err; cut -c1-200 gpurun_out/bench19.json
