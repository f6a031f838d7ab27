cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-140
timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-140
