# scratch: the command list of the last gpurun session (development aid)
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/op_breakdown.py 32 2>&1 | grep -i "cva_residual\|serial step" | head -3
