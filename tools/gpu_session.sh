cd /root/repo
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_kernels.py -x -q -m gpu -k "sda or cross or cva or sample" 2>&1 | tail -2
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, "/root/repo")
from mumpy_b200 import ops
B = 64
pix = torch.rand((B * 64, 3, 49, 2), device="cuda") * 6.0
x2 = torch.randn((B, 3 * 56 * 56, 96), device="cuda")
f = lambda: ops.cva_sample(x2, pix, B, 56, 168, 56, 96, 3, 7, False, torch.bfloat16)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print("cva_sample %.1f us, %.0f GB/s" % (t * 1e3, (6 * x2.numel() + 4 * pix.numel()) / t / 1e6))
PY
