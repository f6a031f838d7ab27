import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
M, C = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn((M, C), device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w1 = (torch.randn((4 * C, C), device=dev) / C ** 0.5).to(dt)
w2 = (torch.randn((C, 4 * C), device=dev) / (4 * C) ** 0.5).to(dt)
b1, b2 = torch.randn(4 * C, device=dev), torch.randn(C, device=dev)
for i in range(2):
    ops.mlp_fused(x, g, b, 1e-5, w1, b1, w2, b2)
    torch.cuda.synchronize()
    print("--", flush=True)
