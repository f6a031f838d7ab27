"""Development aid: CUDA-event timing of one GEMM shape (see tools/profile_kernel.py for the shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops  # noqa: E402

what = sys.argv[1]
B = 32
dev = torch.device("cuda", 0)
shapes = {"gemm_fc1": (B * 588, 2048, 512, ops.ACT_GELU, torch.bfloat16), "gemm_fc2": (B * 588, 512, 2048, ops.ACT_NONE, torch.float32),
          "gemm_s0": (B * 9408, 512, 128, ops.ACT_GELU, torch.bfloat16), "gemm_qkv": (B * 588, 1536, 512, ops.ACT_NONE, torch.bfloat16)}
M, N, K, act, odt = shapes[what]
a = torch.randn((M, K), device=dev).bfloat16()
w = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
bias = torch.zeros(N, device=dev)
res = torch.randn((M, N), device=dev) if odt == torch.float32 else None
out = torch.empty((M, N), dtype=odt, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ts = []
for i in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.linear(a, w, bias, res, act=act, out_dtype=odt, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
t = min(ts[1:])
print("%.1f us  %.0f TFLOP/s" % (t, 2.0 * M * N * K / t / 1e6))
