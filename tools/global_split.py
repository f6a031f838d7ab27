"""Development aid: the 12 temporal ViT blocks (M = 4704 rows at batch 32) as one chain vs k independent chunk chains on k streams
inside one CUDA graph (the blocks act on sequences of 3 tokens: chunks are independent)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda", 0)
enc, dec = bench.build_model(dev)
blocks = enc.base.globalblocks
x = torch.randn((32 * 49, 3, 768), device=dev)
side = [torch.cuda.Stream() for _ in range(4)]


def run(k):
    cur = torch.cuda.current_stream()
    if k == 1:
        return [blocks(x)]
    outs = []
    evs = []
    for i, c in enumerate(x.chunk(k, dim=0)):
        s = side[i]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(blocks(c.contiguous()))
        evs.append(s)
    for s in evs:
        cur.wait_stream(s)
    return outs


with torch.no_grad():
    ref = torch.cat(run(1), 0)
    for k in (1, 2, 3, 4):
        cap = torch.cuda.Stream()
        with torch.cuda.stream(cap):
            o = torch.cat(run(k), 0)
            torch.cuda.synchronize()
            same = bool((o == ref).all())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap):
                keep = run(k)
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("chunks %d on %d streams: %.3f ms for 12 blocks (%.1f us per block), identical: %s" % (k, k, min(ts), min(ts) * 1e3 / 12, same), flush=True)
