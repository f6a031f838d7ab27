cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu21.log | cut -c1-250
timeout 300 python tools/op_breakdown.py 32 > gpurun_out/op_breakdown_v5.txt 2>&1; grep -E "faf|serial step" gpurun_out/op_breakdown_v5.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "bench rc=$?"; tail -3 gpurun_out/bench18.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench18.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e') if k in d})
for k in d.get('kernels',[])[:2]: print(k)
PY
