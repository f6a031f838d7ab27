"""Development aid for ncu: one eager warm-up step then one step bracketed by cudaProfilerStart/Stop.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/profile_step.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mumpy_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
enc, dec = bench.build_model(dev)
x = torch.randn((B, 3, 3, 224, 224), device=dev)
gt = (torch.rand((B, 224, 224), device=dev) > 0.7).to(torch.uint8)
with torch.no_grad():
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        final_x, view_x, ff = enc(x)
        logits, _ = dec(final_x, view_x, ff)
        ops.mask_counts(logits, gt)
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", ops.launch_count)
