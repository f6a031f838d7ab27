cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_gpu24.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu24.log
timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_v10.txt 2>&1; echo "sweep rc=$?"; head -12 gpurun_out/gemm_sweep_v10.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench rc=$?"; cut -c1-180 gpurun_out/bench21.json
