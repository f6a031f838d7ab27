# One launch list of a batch-32 step plus `ncu --set full` summaries of the kernels that matter; the reports are summarised on the
# box (tools/ncu_summary.py) because gpurun brings back at most 64 MiB -- only the three largest kernels' .ncu-rep files are kept.
set -x
mkdir -p gpurun_out/ncu
python tools/profile_step.py 32 > gpurun_out/ncu/plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu/r2_launches_b32.csv python tools/profile_step.py 32 > gpurun_out/ncu/launches.log 2>&1
for k in gemm_tc_kernel ln_gemm_tc_kernel mlp_pipe_tc_kernel window_attention_tc_kernel layernorm_vec_kernel cva_attention_mma resample_rows_kernel groupnorm_apply_rows cva_sample_kernel gather_rows_vec; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/ncu/r2f_$k python tools/profile_step.py 32 > gpurun_out/ncu/$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/ncu/r2f_$k.ncu-rep > gpurun_out/ncu/r2f_${k}_ncu.txt 2>&1
  case $k in gemm_tc_kernel|ln_gemm_tc_kernel|mlp_pipe_tc_kernel) ;; *) rm -f gpurun_out/ncu/r2f_$k.ncu-rep ;; esac
done
du -sh gpurun_out
