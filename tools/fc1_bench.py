"""Development aid: the GELU GEMMs (fc1, 16-bit output) of a B=32 step in isolation, plain and LayerNorm-fused."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


for M, N, K in [(18816, 2048, 512), (6272, 1536, 384), (4704, 3072, 768), (1568, 4096, 1024), (75264, 1024, 256), (301056, 512, 128)]:
    a = torch.randn((M, K), device=dev).to(dt)
    x = torch.randn((M, K), device=dev)
    g, b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    w = (torch.randn((N, K), device=dev) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=dev)
    t = timed(lambda: ops.linear(a, w, bias, act=ops.ACT_GELU, out_dtype=dt))
    t0 = timed(lambda: ops.linear(a, w, bias, out_dtype=dt))
    line = "M=%6d N=%4d K=%4d | GELU %6.1f us (%4.0f TFLOP/s) | no act %6.1f us" % (M, N, K, t, 2.0 * M * N * K / t * 1e-6, t0)
    if ops.ln_linear_fits(N, K):
        t2 = timed(lambda: ops.ln_linear(x, g, b, 1e-5, w, bias, act=ops.ACT_GELU))
        line += " | LN-fused GELU %6.1f us" % t2
    print(line, flush=True)
