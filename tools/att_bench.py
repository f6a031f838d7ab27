"""Times table-mode window attention (16-bit) for the model's shapes at B=32 with both kernels (tcgen05 vs mma.sync)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mumpy_b200
from mumpy_b200 import ops

SHAPES = [  # (T*H, W, C, heads) per view/stage at 224^2, B=32
    (168, 56, 128, 4), (56, 56, 96, 3), (84, 28, 256, 8), (28, 28, 192, 6), (42, 14, 512, 16), (14, 14, 384, 12), (21, 7, 1024, 32), (7, 7, 768, 24)]
B = 32
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


ops.set_precision("bf16")
for TH, W, C, heads in SHAPES:
    for shift in (0, 3):
        if min(TH, W) <= 7 and shift:
            continue
        qkv = torch.randn(B, TH * W, 3 * C, device=dev).bfloat16()
        table = torch.randn(169, heads, device=dev) * 0.5
        from oracle import mumpy_oracle as orc
        bias = orc.relative_position_bias(table.cpu(), 7).to(dev)
        mask = orc.shifted_window_mask(TH, W, 7, shift).to(dev) if shift else None
        res = {}
        outs = {}
        for tc in (0, 1):
            ops.set_attention_tc(bool(tc))
            f = lambda: ops.window_attention(qkv, bias, mask, B, TH, W, C, heads, 7, shift, rel_table=table, standard_mask=mask is not None)
            outs[tc] = f().float()
            res[tc] = timed(f)
        byts = qkv.numel() * 2 + qkv.numel() // 3 * 2
        diff = (outs[0] - outs[1]).abs().max().item()
        print("TH=%3d W=%2d C=%4d heads=%2d shift=%d  mma %7.1f us  tc %7.1f us  (%.2fx)  tc %.0f GB/s  maxdiff %.4f" % (
            TH, W, C, heads, shift, res[0], res[1], res[0] / res[1], byts / res[1] / 1e3, diff), flush=True)
