cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu16.log
timeout 300 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_v9.txt 2>&1; echo "sweep rc=$?"; head -12 gpurun_out/gemm_sweep_v9.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-kernels --no-fp16 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"; tail -3 gpurun_out/bench15.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench15.json'))
print({k:d[k] for k in ('value','ms_per_step','roofline') if k in d})
PY
