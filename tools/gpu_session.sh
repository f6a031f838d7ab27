cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py tests/test_gpu_e2e.py -m gpu -x -q -s > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?"
grep -E "mask identity|passed|failed|Error|error" gpurun_out/pytest_gpu15.log | head -20
timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_v8.txt 2>&1; echo "sweep rc=$?"; head -12 gpurun_out/gemm_sweep_v8.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"; tail -3 gpurun_out/bench14.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench14.json'))
print({k:d[k] for k in ('value','ms_per_step','roofline') if k in d})
PY
