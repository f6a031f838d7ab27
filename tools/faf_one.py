import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mumpy_b200
from mumpy_b200.models.modules.dct import FAF
dev = torch.device("cuda", 0)
faf = FAF(224).eval()
for B in (32, 64):
    x = torch.randn((B, 3, 3, 224, 224), device=dev)
    faf.frame(x, 1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        torch.cuda._sleep(2000000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        faf.frame(x, 1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    byts = B * 224 * 224 * (3 * 4 + 9 * 4)
    print("faf16 B=%d: %.1f us, %.0f GB/s algorithmic (%.1f %% of 6539)" % (B, min(ts), byts / min(ts) * 1e-3, byts / min(ts) * 1e-3 / 65.39), flush=True)
