# Development aid: same-box A/B of two builds of the library (tools/ab/libmumpy_b200_base.so vs the in-tree one), alternating.
for i in 1 2 3; do
  for v in base new; do
    if [ $v = base ]; then export MUMPY_LIB=$PWD/tools/ab/libmumpy_b200_base.so; else unset MUMPY_LIB; fi
    timeout 300 python bench.py --no-kernels --no-eager --no-split --no-fp16 --cpu-clips 2 > gpurun_out/ab_bench.json 2>gpurun_out/ab_bench.err
    python -c "
import json
d=json.loads(open('gpurun_out/ab_bench.json').read().strip().splitlines()[-1]); print('$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['serial_kernel_time_ms'],3), d['clocks']['sm_mhz'])
"
  done
done
