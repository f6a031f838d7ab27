set -x
cd /root/repo
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json; cut -c1-200 gpurun_out/bench_final.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_frontend.py tests/test_gpu_kernels.py -x -q -m gpu -k "resize or window_attention or split_k or native" 2>&1 | tail -12
