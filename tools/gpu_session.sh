# scratch: the command list of the last gpurun session (development aid)
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-160
