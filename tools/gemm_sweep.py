"""Development aid: records every GEMM / conv shape of one B-clip forward, then times each distinct shape in isolation
(CUDA events, L2 flushed) and prints its share, TFLOP/s and GB/s."""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mumpy_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
enc, dec = bench.build_model(dev)
x = torch.randn((B, 3, 3, 224, 224), device=dev)
log = OrderedDict()
orig_linear, orig_conv = ops.linear, ops.conv2d_nhwc_bf16


def rec_linear(a, w, bias=None, residual=None, act=ops.ACT_NONE, out_dtype=torch.float32, out=None):
    K = a.shape[-1]
    key = ("lin", a.numel() // K, w.shape[0], K, act, str(out_dtype), residual is not None)
    log[key] = log.get(key, 0) + 1
    return orig_linear(a, w, bias, residual, act, out_dtype, out)


def rec_conv(x, wq, bias, B_, H, W, Cin, Cout, kh, kw, ph, pw, ld_in=None, act=ops.ACT_NONE, out_dtype=torch.float32, residual=None):
    key = ("conv", B_, H, W, Cin, Cout, kh, kw)
    log[key] = log.get(key, 0) + 1
    return orig_conv(x, wq, bias, B_, H, W, Cin, Cout, kh, kw, ph, pw, ld_in, act, out_dtype, residual)


ops.linear, ops.conv2d_nhwc_bf16 = rec_linear, rec_conv
with torch.no_grad():
    f, v, ff = enc(x)
    dec(f, v, ff)
torch.cuda.synchronize()
ops.linear, ops.conv2d_nhwc_bf16 = orig_linear, orig_conv
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)


rows = []
for key, cnt in log.items():
    if key[0] == "lin":
        _, M, N, K, act, odt, has_res = key
        odt = torch.bfloat16 if "bfloat16" in odt else torch.float32
        a = torch.randn((M, K), device=dev).bfloat16()
        w = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
        bias = torch.zeros(N, device=dev)
        res = torch.randn((M, N), device=dev) if has_res else None
        t = timed(lambda: orig_linear(a, w, bias, res, act, odt))
        flops = 2.0 * M * N * K
        byts = 2 * M * K + 2 * N * K + (2 if odt == torch.bfloat16 else 4) * M * N + (4 * M * N if has_res else 0)
        name = "lin M=%d N=%d K=%d act=%d %s%s" % (M, N, K, act, "bf16" if odt == torch.bfloat16 else "f32", "+res" if has_res else "")
    else:
        _, B_, H, W, Cin, Cout, kh, kw = key
        ld = (Cin + 7) // 8 * 8
        xin = torch.randn((B_, H, W, ld), device=dev).bfloat16()
        cb = (Cin + 63) // 64
        wq = (torch.randn((Cout, kh * kw * cb * 64), device=dev) / (Cin * kh * kw) ** 0.5).bfloat16()
        bias = torch.zeros(Cout, device=dev)
        t = timed(lambda: orig_conv(xin, wq, bias, B_, H, W, Cin, Cout, kh, kw, (kh - 1) // 2, (kw - 1) // 2, ld_in=ld))
        M = B_ * H * W
        flops = 2.0 * M * Cout * Cin * kh * kw
        byts = 2 * M * Cin + 2 * wq.numel() + 4 * M * Cout
        name = "conv %dx%d %dx%dx%d Cin=%d Cout=%d" % (kh, kw, B_, H, W, Cin, Cout)
    rows.append((t * cnt, cnt, t, flops / t / 1e12, byts / t / 1e9, name))
tot = sum(r[0] for r in rows)
print("total GEMM+conv time per step (isolated, cold L2): %.2f ms" % (tot * 1e3))
for tt, cnt, t, tf, gb, name in sorted(rows, reverse=True):
    print("%6.1f us x%3d = %7.1f us (%4.1f%%)  %6.1f TFLOP/s %6.0f GB/s  %s" % (t * 1e6, cnt, tt * 1e6, 100 * tt / tot, tf, gb, name))
