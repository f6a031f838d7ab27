import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
M, N, K, gelu = [int(v) for v in sys.argv[1:5]]
x = torch.randn((M, K), device=dev)
g, b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
w = (torch.randn((N, K), device=dev) / K ** 0.5).to(dt)
bias = torch.randn(N, device=dev)
for i in range(3):
    ops.ln_linear(x, g, b, 1e-5, w, bias, act=ops.ACT_GELU if gelu else ops.ACT_NONE)
    torch.cuda.synchronize()
    print("--", flush=True)
