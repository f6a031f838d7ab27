cd /root/repo
echo REG; timeout 120 python tools/cva_one.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3
