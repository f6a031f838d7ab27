"""Development aid: in-situ time per library op of one eager forward on a single stream (bench.timed_serial_step: CUDA events
around every ops.* call behind a parked stream, warm L2 as in the real step), aggregated by op and by (op, shape).
    python tools/op_breakdown.py [B]"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
enc, dec = bench.build_model(dev)
x = torch.randn((B, 3, 3, 224, 224), device=dev)
gt = (torch.rand((B, 224, 224), device=dev) > 0.7).to(torch.uint8)
with torch.no_grad():
    recs = bench.timed_serial_step(enc, dec, x, gt)
by_op, by_shape = defaultdict(lambda: [0.0, 0]), defaultdict(lambda: [0.0, 0])
for name, shp, _, t in recs:
    t *= 1e6
    by_op[name][0] += t
    by_op[name][1] += 1
    by_shape[(name, shp)][0] += t
    by_shape[(name, shp)][1] += 1
acc = sum(v[0] for v in by_op.values())
print("serial step: %.0f us inside library ops over %d calls (B=%d)" % (acc, len(recs), B))
for n, (t, c) in sorted(by_op.items(), key=lambda kv: -kv[1][0]):
    print("%9.1f us %5.1f%% %4d x  %s" % (t, 100 * t / acc, c, n))
print("\ntop (op, shapes):")
for (n, shp), (t, c) in sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%9.1f us %4d x %7.1f us  %s %s" % (t, c, t / c, n, shp))
