# scratch: the command list of the last gpurun session (development aid)
set -x
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --no-kernels 2>&1 | tail -1 | cut -c1-200
