set -x
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_line.json; cut -c1-300 gpurun_out/bench_line.json
