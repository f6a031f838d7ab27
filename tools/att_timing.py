import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mumpy_b200 import ops
from oracle import mumpy_oracle as orc
B, TH, W, C, heads = 32, 42, 14, 512, 16
if len(sys.argv) > 1:
    B, TH, W, C, heads = [int(v) for v in sys.argv[1:6]]
qkv = torch.randn(B, TH * W, 3 * C, device="cuda").half()
table = torch.randn(169, heads, device="cuda") * 0.5
bias = orc.relative_position_bias(table.cpu(), 7).cuda()
mask = orc.shifted_window_mask(TH, W, 7, 3).cuda()
for shift in (0, 3):
    for _ in range(2):
        ops.window_attention(qkv, bias, mask if shift else None, B, TH, W, C, heads, 7, shift, rel_table=table, standard_mask=bool(shift))
        torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        torch.cuda._sleep(600000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.window_attention(qkv, bias, mask if shift else None, B, TH, W, C, heads, 7, shift, rel_table=table, standard_mask=bool(shift))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    byts = B * TH * W * C * 2 * 4
    print("window attention B=%d TH=%d W=%d C=%d heads=%d shift=%d: %.1f us, %.0f GB/s (%.1f %% of 6539)" % (B, TH, W, C, heads, shift, min(ts), byts / min(ts) * 1e-3, byts / min(ts) * 1e-3 / 65.39), flush=True)
