"""Development aid: mumpy_ln_linear against mumpy_layernorm + mumpy_linear on the Swin shapes of a B=32 step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
SHAPES = [(18816, 1536, 512, 0), (18816, 2048, 512, 1), (6272, 1152, 384, 0), (6272, 1536, 384, 1), (75264, 768, 256, 0), (75264, 1024, 256, 1),
          (25088, 576, 192, 0), (25088, 768, 192, 1), (301056, 384, 128, 0), (301056, 512, 128, 1), (100352, 288, 96, 0), (100352, 384, 96, 1)]
if len(sys.argv) > 2:
    SHAPES = SHAPES[int(sys.argv[1]):int(sys.argv[2])]
elif len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=5, do_flush=True):
    """min over reps of the device time of fn(); the stream is parked behind a spin kernel while the host enqueues, so the
    interval holds no launch latency.  do_flush: L2 is flushed first (cold operands), else the previous rep left them warm."""
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        torch.cuda._sleep(600000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


dt = torch.float16
for M, N, K, gelu in SHAPES:
    x = torch.randn((M, K), device=dev)
    g, b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    w = (torch.randn((N, K), device=dev) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=dev)
    act = ops.ACT_GELU if gelu else ops.ACT_NONE
    xn = ops.layernorm(x, g, b, 1e-5, out_dtype=dt)
    res = {}
    for cold in (True, False):
        t_ln = timed(lambda: ops.layernorm(x, g, b, 1e-5, out_dtype=dt), do_flush=cold)
        t_mm = timed(lambda: ops.linear(xn, w, bias, act=act, out_dtype=dt), do_flush=cold)
        t_un = timed(lambda: ops.linear(ops.layernorm(x, g, b, 1e-5, out_dtype=dt), w, bias, act=act, out_dtype=dt), do_flush=cold)
        t_f = timed(lambda: ops.ln_linear(x, g, b, 1e-5, w, bias, act=act), do_flush=cold)
        res[cold] = (t_ln, t_mm, t_un, t_f)
    fl = 2.0 * M * N * K
    print("M=%6d N=%4d K=%3d g%d | cold: ln %5.1f gemm %5.1f (%4.0f TF) ln+gemm %5.1f fused %5.1f (%4.0f TF) x%.2f | warm: ln %5.1f gemm %5.1f (%4.0f TF) ln+gemm %5.1f fused %5.1f (%4.0f TF) x%.2f"
          % ((M, N, K, gelu) + tuple(v for c in (True, False) for v in (res[c][0], res[c][1], fl / res[c][1] * 1e-6, res[c][2], res[c][3], fl / res[c][3] * 1e-6, res[c][2] / res[c][3]))), flush=True)
