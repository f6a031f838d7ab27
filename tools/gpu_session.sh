set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu10.log
for cfg in "MUMPY_PDL=1 MUMPY_STREAMS=1" "MUMPY_PDL=0 MUMPY_STREAMS=1" "MUMPY_PDL=0 MUMPY_STREAMS=0"; do
  tag=$(echo $cfg | tr -d ' =A-Z_')
  env $cfg python bench.py --steps 5 --warmup 3 --no-kernels > gpurun_out/bench7_$tag.json 2> gpurun_out/bench7_$tag.err; echo "bench $cfg rc=$?"
  cut -c1-200 gpurun_out/bench7_$tag.json; tail -2 gpurun_out/bench7_$tag.err
  env $cfg python bench.py --steps 5 --warmup 3 --no-kernels --batch 1 > gpurun_out/bench7_b1_$tag.json 2>/dev/null; cut -c60-200 gpurun_out/bench7_b1_$tag.json
done
