cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu17.log
timeout 300 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown_v4.txt 2>&1; grep -E "tokenize|serial step" gpurun_out/op_breakdown_v4.txt
for cfg in "MUMPY_PDL=1" "MUMPY_PDL=0"; do
timeout 600 env $cfg python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench16_$cfg.json 2> gpurun_out/bench16.err; echo "bench $cfg rc=$?"; cut -c1-180 gpurun_out/bench16_$cfg.json
done
