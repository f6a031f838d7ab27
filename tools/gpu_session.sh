cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --size 512 --batch 8 --steps 10 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench512 rc=$?"; tail -3 gpurun_out/bench_512.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_512.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','other_precision','roofline','config') if k in d})
PY
