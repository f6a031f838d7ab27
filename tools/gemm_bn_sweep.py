"""Development aid: for the step's heaviest GEMM shapes, time every legal tile width (mumpy_set_gemm_tile) against the cost
model's choice, cold L2.   python tools/gemm_bn_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
SHAPES = [  # (M, N, K, act, out16, residual)
    (18816, 2048, 512, 1, True, False), (18816, 512, 2048, 0, False, True), (18816, 1536, 512, 0, True, False), (18816, 512, 512, 0, False, True),
    (6272, 1536, 384, 1, True, False), (6272, 384, 1536, 0, False, True), (6272, 1152, 384, 0, True, False), (6272, 384, 384, 0, False, True),
    (4704, 3072, 768, 1, True, False), (4704, 768, 3072, 0, False, True), (4704, 2304, 768, 0, True, False), (4704, 768, 768, 0, False, True),
    (301056, 512, 128, 1, True, False), (301056, 128, 512, 0, False, True), (301056, 384, 128, 0, True, False), (301056, 128, 128, 0, False, True),
    (100352, 384, 96, 1, True, False), (100352, 96, 384, 0, False, True), (100352, 288, 96, 0, True, False),
    (75264, 1024, 256, 1, True, False), (75264, 256, 1024, 0, False, True), (75264, 768, 256, 0, True, False),
    (25088, 768, 192, 1, True, False), (25088, 192, 768, 0, False, True), (1568, 3072, 768, 1, True, False), (1568, 768, 3072, 0, False, True),
]


def timed(fn, reps=5):
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
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


total_auto = total_best = 0.0
for M, N, K, act, out16, res in SHAPES:
    a = torch.randn((M, K), device=dev).bfloat16()
    w = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
    bias = torch.zeros(N, device=dev)
    r = torch.randn((M, N), device=dev) if res else None
    odt = torch.bfloat16 if out16 else torch.float32
    out = torch.empty((M, N), dtype=odt, device=dev)
    row = []
    lib.mumpy_set_gemm_tile(0)
    ops.set_gemm_pair_mode(1)
    t_auto = timed(lambda: ops.linear(a, w, bias, r, act=act, out_dtype=odt, out=out))
    for pm in (0, 2):
        ops.set_gemm_pair_mode(pm)
        for bn in (256, 192, 128, 96, 64, 48, 32):
            if N % bn:
                continue
            lib.mumpy_set_gemm_tile(bn)
            row.append((timed(lambda: ops.linear(a, w, bias, r, act=act, out_dtype=odt, out=out)), bn, pm))
    lib.mumpy_set_gemm_tile(0)
    ops.set_gemm_pair_mode(1)
    best = min(row)
    total_auto += t_auto
    total_best += best[0]
    print("M=%6d N=%4d K=%4d act=%d %s%s  auto %6.1f us | best BN=%3d%s %6.1f us | 1-CTA %s | pair %s" % (
        M, N, K, act, "16" if out16 else "32", "+res" if res else "    ", t_auto, best[1], "p" if best[2] else " ", best[0],
        "  ".join("%d:%.1f" % (bn, t) for t, bn, pm in sorted(row, key=lambda x: -x[1]) if pm == 0),
        "  ".join("%d:%.1f" % (bn, t) for t, bn, pm in sorted(row, key=lambda x: -x[1]) if pm == 2)), flush=True)
print("sum over the shapes: cost model %.1f us, per-shape optimum %.1f us" % (total_auto, total_best))
