cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/op_breakdown.py 32 2>&1 | grep -i "groupnorm_nhwc\|resample\|gather_rows\|serial step" | head -8
timeout 600 python bench.py --no-kernels --no-fp16 2>&1 | tail -1 | cut -c1-140
