"""Development aid: where one B=32 forward spends its time on the critical path -- CUDA events on the caller's stream at the
stage boundaries of an eager forward with the lanes on (each boundary joins the lanes it depends on)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mumpy_b200
from mumpy_b200 import ops, streams
from mumpy_b200.models.encoder import multiTemporalViewEncoder as mtv
dev = torch.device("cuda", 0)
enc, dec = bench.build_model(dev)
x = bench.synthetic_batches(32, 224, rank=0, n=1)[0].to(dev)
marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


base = enc.base
orig_layers = [l.forward for l in base.layers.layers]
for i, lyr in enumerate(base.layers.layers):
    def f(xs, _o=orig_layers[i], _i=i):
        r = _o(xs)
        # join the three lanes on the current stream for the measurement only
        torch.cuda.synchronize()
        mark("stage %d" % _i)
        return r
    lyr.forward = f
orig_tok = base.tokenize.forward
def tok(xx):
    r = orig_tok(xx)
    torch.cuda.synchronize()
    mark("tokenize (+faf lane)")
    return r
base.tokenize.forward = tok
orig_glob = base._global_part
def glob(xs, B):
    r = orig_glob(xs, B)
    torch.cuda.synchronize()
    mark("global blocks")
    return r
base._global_part = glob
with torch.no_grad():
    for it in range(3):
        del marks[:]
        torch.cuda.synchronize()
        mark("start")
        out = mumpy_b200.forward(enc, dec, x)
        torch.cuda.synchronize()
        mark("decoder")
prev = marks[0][1]
tot = 0.0
for name, e in marks[1:]:
    dt = prev.elapsed_time(e)
    tot += dt
    print("%-24s %7.3f ms" % (name, dt))
    prev = e
print("%-24s %7.3f ms (eager, synchronised at every boundary)" % ("total", tot))
