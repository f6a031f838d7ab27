cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu30.log | cut -c1-200
timeout 300 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown_v10.txt 2>&1; grep -E "groupnorm|serial step" gpurun_out/op_breakdown_v10.txt | head
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench26.json 2> gpurun_out/bench26.err; echo "bench rc=$?"; cut -c1-180 gpurun_out/bench26.json
