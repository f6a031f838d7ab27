"""Development aid: mumpy_cva_attention alone on the cross-view shapes of a batch-32 step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for B, TH1, TH2, W, C, heads in [(32, 56, 168, 56, 96, 3), (32, 56, 56, 56, 96, 3), (32, 28, 84, 28, 192, 6), (32, 14, 42, 14, 384, 12), (32, 7, 21, 7, 768, 24)]:
    ws = 7
    N2 = B * (TH2 // ws) * (W // ws)
    q = torch.randn((B, TH1 * W, C), device=dev)
    kv = torch.randn((N2 * 49, 2 * C), device=dev).to(dt)
    fn = lambda: ops.cva_attention(q, kv, B, TH1, TH2, W, C, heads, ws, False)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    byts = q.numel() * 4 * (TH2 // TH1) + kv.numel() * 2 + q.numel() * 2
    print("cva_attention B=%d TH1=%d TH2=%d W=%d C=%d heads=%d: %.1f us (%.0f GB/s)" % (B, TH1, TH2, W, C, heads, min(ts), byts / min(ts) * 1e-3), flush=True)
