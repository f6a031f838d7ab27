import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
SHAPES = [(18816, 512, 2048, 0, 0, 1), (18816, 512, 512, 0, 0, 1), (6272, 384, 1536, 0, 0, 1), (6272, 384, 384, 0, 0, 1), (4704, 768, 3072, 0, 0, 1), (4704, 768, 768, 0, 0, 1),
          (301056, 128, 512, 0, 0, 1), (75264, 256, 1024, 0, 0, 1), (18816, 1536, 512, 0, 1, 0), (18816, 2048, 512, 1, 1, 0), (4704, 3072, 768, 1, 1, 0), (4704, 2304, 768, 0, 1, 0)]


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda._sleep(600000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


for M, N, K, gelu, out16, res in SHAPES:
    a = torch.randn((M, K), device=dev).to(dt)
    w = (torch.randn((N, K), device=dev) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=dev)
    r = torch.randn((M, N), device=dev) if res else None
    out = torch.empty((M, N), device=dev, dtype=dt if out16 else torch.float32)
    row = []
    for pm in (0, 2):
        ops.set_gemm_pair_mode(pm)
        t = timed(lambda: ops.linear(a, w, bias, r, act=ops.ACT_GELU if gelu else ops.ACT_NONE, out_dtype=out.dtype, out=out))
        row.append(t)
    ops.set_gemm_pair_mode(0)
    fl = 2.0 * M * N * K
    print("M=%6d N=%4d K=%4d gelu=%d out16=%d res=%d | 1-CTA %6.1f us (%4.0f TF)  pair %6.1f us (%4.0f TF)  x%.2f" % (M, N, K, gelu, out16, res, row[0], fl / row[0] * 1e-6, row[1], fl / row[1] * 1e-6, row[0] / row[1]), flush=True)
