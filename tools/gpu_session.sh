cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"
grep -E "mask identity|passed|failed|Error|error" gpurun_out/pytest_gpu13.log | head -20; tail -5 gpurun_out/pytest_gpu13.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-kernels > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench11.json
