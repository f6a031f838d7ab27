# Development aid: same-box A/B of an environment knob: bash tools/ab_env.sh NAME VALUE_A VALUE_B
run() {
  timeout 300 python bench.py --no-kernels --no-eager --no-split --no-fp16 --cpu-clips 2 > gpurun_out/ab_bench.json 2>gpurun_out/ab_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab_bench.json').read().strip().splitlines()[-1]); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['serial_kernel_time_ms'],3), d['clocks']['sm_mhz'])
"
}
for i in 1 2; do
  for v in "$2" "$3"; do
    export $1=$v
    run "$1=$v"
  done
done
