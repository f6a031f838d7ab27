cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu25.log
timeout 300 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown_v6.txt 2>&1; grep -E "window_attention|serial step" gpurun_out/op_breakdown_v6.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench22.json 2> gpurun_out/bench22.err; echo "bench rc=$?"; cut -c1-180 gpurun_out/bench22.json
