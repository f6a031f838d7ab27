cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu31.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu31.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench27.json 2> gpurun_out/bench27.err; echo "bench rc=$?"; cut -c1-180 gpurun_out/bench27.json; tail -3 gpurun_out/bench27.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 --batch 1 > gpurun_out/bench27_b1.json 2> /dev/null; cut -c60-180 gpurun_out/bench27_b1.json
