"""Development aid: stage-2 view-3 Swin blocks 1..17 on the full B=32 canvas vs the batch cut into chunks that run the
whole chain one after the other (the chunk's activations then stay in L2 between producer and consumer kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mumpy_b200 import ops
dev = torch.device("cuda", 0)
enc, dec = bench.build_model(dev)
view = int(sys.argv[1]) if len(sys.argv) > 1 else 3
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 2
blocks = [getattr(b, "block%d" % view) for b in enc.base.layers.layers[stage].blocks[1:]]
blocks = [b for b in blocks if not isinstance(b, torch.nn.Identity)]
res = blocks[0].input_resolution
C = blocks[0].dim
T = {3: 3, 2: 1, 1: 1}[view]
B = 32
L = T * res[0] * res[1]
x = torch.randn((B, L, C), device=dev)
print("view %d stage %d: %d blocks, canvas (%d, %d, %d)" % (view, stage, len(blocks), B, L, C))


def chain(t):
    for b in blocks:
        t = b(t)
    return t


def run(nchunk):
    outs = []
    for c in x.chunk(nchunk, dim=0):
        outs.append(chain(c))
    return outs


with torch.no_grad():
    ref = torch.cat(run(1), 0)
    for n in (1, 2, 4):
        o = torch.cat(run(n), 0)
        same = bool((o == ref).all())
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            run(n)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                run(n)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("chunks %d: %.3f ms per chain (%.1f us per block), identical to unchunked: %s" % (n, min(ts), min(ts) * 1e3 / len(blocks), same), flush=True)
