cd $GRAFT_REPO_ROOT
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print(json.dumps({k:d[k] for k in d if k not in ('kernels',)}, indent=None)[:3000])
for k in d.get('kernels',[]): print(k)
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
