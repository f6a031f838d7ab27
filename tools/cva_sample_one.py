import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mumpy_b200 import ops
B = 64
pix = torch.rand((B * 64, 3, 49, 2), device="cuda") * 6.0
x2 = torch.randn((B, 3 * 56 * 56, 96), device="cuda")
for _ in range(3):
    ops.cva_sample(x2, pix, B, 56, 168, 56, 96, 3, 7, False, torch.bfloat16)
    torch.cuda.synchronize()
