"""Development aid: mumpy_mlp_fused against LayerNorm + fc1 + fc2 on the stage-0/1 shapes of a B=32 step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda._sleep(1000000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


for M, C in [(301056, 128), (75264, 256), (100352, 96), (25088, 192)]:
    x = torch.randn((M, C), device=dev)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    w1 = (torch.randn((4 * C, C), device=dev) / C ** 0.5).to(dt)
    w2 = (torch.randn((C, 4 * C), device=dev) / (4 * C) ** 0.5).to(dt)
    b1, b2 = torch.randn(4 * C, device=dev), torch.randn(C, device=dev)
    t_un = timed(lambda: ops.linear(ops.linear(ops.layernorm(x, g, b, 1e-5, out_dtype=dt), w1, b1, act=ops.ACT_GELU, out_dtype=dt), w2, b2, residual=x))
    t_f = timed(lambda: ops.mlp_fused(x, g, b, 1e-5, w1, b1, w2, b2))
    fl = 2.0 * M * C * 4 * C * 2
    byts = M * C * 4 * 2
    print("M=%6d C=%3d | LN+fc1+fc2 %6.1f us | fused %6.1f us (%4.0f TFLOP/s, %4.0f GB/s algorithmic) x%.2f" % (M, C, t_un, t_f, fl / t_f * 1e-6, byts / t_f * 1e-3, t_un / t_f), flush=True)
