set -x
cd /root/repo
timeout 600 python bench.py --no-kernels --no-fp16 --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nofp16', d['value'], d['roofline']['share_of_kernel_time'], d['roofline']['serial_kernel_time_ms'], d['roofline']['achieved'])"
timeout 600 python bench.py --no-kernels --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('withfp16', d['value'], d['roofline']['share_of_kernel_time'], d['roofline']['serial_kernel_time_ms'], d['roofline']['achieved'])"
