"""Development aid: does alternating between big-shared-memory GEMM kernels and small elementwise kernels cost anything per switch
(shared-memory carve-out reconfiguration)?  Times CUDA graphs of the same 40 kernels in alternating and in grouped order."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
dt = torch.float16
M, K, N = 6272, 384, 384
x = torch.randn((M, K), device=dev)
g, b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
w = (torch.randn((N, K), device=dev) / K ** 0.5).to(dt)
xn = ops.layernorm(x, g, b, 1e-5, out_dtype=dt)
out = torch.empty((M, N), device=dev)
xo = torch.empty_like(xn)


def ln():
    lib, st = ops._prep(x, g, b, xo)
    ops._lib.check(lib.mumpy_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), xo.data_ptr(), ops.code(dt), M, K, 1e-5, st), "ln")


def mm():
    ops.linear(xn, w, None, None, out=out)


def graph_of(seq):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in seq:
            f()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for f in seq:
                f()
    return gr


def timed(gr, reps=20):
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


n = 20
alt = graph_of([f for _ in range(n) for f in (mm, ln)])
grp = graph_of([mm] * n + [ln] * n)
only_mm = graph_of([mm] * n)
only_ln = graph_of([ln] * n)
for name, gr in (("alternating", alt), ("grouped", grp), ("gemm only", only_mm), ("ln only", only_ln)):
    print("%-12s %7.1f us per graph" % (name, timed(gr)), flush=True)
