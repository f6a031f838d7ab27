set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu8.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench5.json
python tools/gemm_sweep.py > gpurun_out/gemm_sweep_v3.txt 2>&1; echo "sweep rc=$?"
python tools/profile_step.py 32 > gpurun_out/prof_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_b32_v4.csv python tools/profile_step.py 32 > gpurun_out/prof_ncu.log 2>&1; echo "launches rc=$?"
python tools/profile_kernel.py gemm_small > gpurun_out/pk_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/prof_gemm_small_v4 -f python tools/profile_kernel.py gemm_small > gpurun_out/pk_ncu.log 2>&1; echo "ncu gemm rc=$?"
python tools/profile_kernel.py attn > gpurun_out/pk_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:window_attention -s 1 -c 1 -o gpurun_out/prof_attn_v4 -f python tools/profile_kernel.py attn > gpurun_out/pk_ncu2.log 2>&1; echo "ncu attn rc=$?"
