import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mumpy_b200 import ops
from oracle import mumpy_oracle as orc
ops.set_precision("bf16")
B, TH, W, C, heads = 32, 168, 56, 128, 4
qkv = torch.randn(B, TH * W, 3 * C, device="cuda").bfloat16()
table = torch.randn(169, heads, device="cuda") * 0.5
bias = orc.relative_position_bias(table.cpu(), 7).cuda()
for _ in range(2):
    ops.window_attention(qkv, bias, None, B, TH, W, C, heads, 7, 0, rel_table=table, standard_mask=False)
    torch.cuda.synchronize()
