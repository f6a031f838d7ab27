# Development aid: fused-MLP width policies on one box (base = the committed baseline library under tools/ab/)
run() {
  timeout 300 python bench.py --no-kernels --no-eager --no-split --no-fp16 --cpu-clips 2 > gpurun_out/ab_bench.json 2>gpurun_out/ab_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_bench.json').read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['serial_kernel_time_ms'],3), d['clocks']['sm_mhz'])
"
}
for i in 1 2; do
  MUMPY_LIB=$PWD/tools/ab/libmumpy_b200_base.so run base
  MUMPY_FUSED_MLP_WIDTHS=96,128,192,256 run all
  MUMPY_FUSED_MLP_WIDTHS=96,128 run narrow
  MUMPY_FUSED_MLP_WIDTHS=96,128,256 run no192
done
