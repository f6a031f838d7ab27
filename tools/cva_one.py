"""Times mumpy_cva_offsets for the four stage shapes (B=32); MUMPY_CVA_REG=0 selects the CTA-per-unit kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from mumpy_b200 import ops
B, groups, ws = 32, 3, 7
for TH1, C in ((56, 96), (28, 192), (14, 384), (7, 768)):
    W, Cg = TH1, C // groups
    q = torch.randn(B, TH1 * W, C, device="cuda")
    dw_w = torch.randn(Cg, 1, 5, 5, device="cuda") * 0.2
    dw_b, ln_g, ln_b = torch.randn(Cg, device="cuda"), torch.ones(Cg, device="cuda"), torch.zeros(Cg, device="cuda")
    pw = torch.randn(2, Cg, device="cuda") * 0.1
    f = lambda: ops.cva_offsets(q, dw_w.view(-1), dw_b, ln_g, ln_b, pw.view(-1), B, TH1, W, C, groups, ws)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record(); torch.cuda.synchronize()
    print("cva_offsets TH1=%d C=%d Cg=%d units=%d: %.1f us" % (TH1, C, Cg, B * (TH1 // 7) ** 2 * 3, e0.elapsed_time(e1) * 100))
