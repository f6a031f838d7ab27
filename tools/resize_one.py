import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mumpy_b200 import ops
fr = torch.randint(0, 256, (64, 480, 854, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    ops.resize_u8(fr, 224, 224)
    torch.cuda.synchronize()
