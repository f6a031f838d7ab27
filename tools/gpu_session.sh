set -x
cd /root/repo
timeout 300 python tools/att_bench.py 2>&1 | tail -20
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --no-kernels 2>&1 | tail -1 | cut -c1-400
