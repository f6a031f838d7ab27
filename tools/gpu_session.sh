cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_e2e.py -x -q -m gpu -k "cout1 or fp32_mode or decoder" 2>&1 | tail -3
timeout 600 python tools/op_breakdown.py 32 2>&1 | grep -i "cout1\|serial step"
