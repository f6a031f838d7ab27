import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
for M, N, K, gelu, out16, res in [(18816, 1536, 512, 0, 1, 0), (75264, 1024, 256, 1, 1, 0)]:
    a = torch.randn((M, K), device=dev).to(dt)
    w = (torch.randn((N, K), device=dev) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=dev)
    r = torch.randn((M, N), device=dev) if res else None
    for i in range(2):
        ops.linear(a, w, bias, r, act=ops.ACT_GELU if gelu else ops.ACT_NONE, out_dtype=dt if out16 else torch.float32)
        torch.cuda.synchronize()
    print("----", flush=True)
