cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu18.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"; tail -3 gpurun_out/bench17.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench17.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','other_precision','gpu_launches','clocks') if k in d})
for k in d.get('kernels',[]): print(k)
PY
