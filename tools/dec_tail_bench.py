"""Development aid: the decoder tail at batch 32 -- final_out conv (Cout = 1), GroupNorm at 112^2, gated upsample."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops
dev = torch.device("cuda", 0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda._sleep(1000000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


B = 32
x = torch.randn((B, 224, 224, 32), device=dev)
w = torch.randn((1, 3, 3, 32), device=dev)
bias = torch.randn(1, device=dev)
t = timed(lambda: ops.conv2d_nhwc_cout1(x, w, bias, B, 224, 224, 32, 3, 3, 1, 1))
print("conv_cout1 (32,224,224,32): %.1f us, %.0f GB/s" % (t, x.numel() * 4 / t * 1e-3), flush=True)
y = torch.randn((B, 112, 112, 128), device=dev)
g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
t = timed(lambda: ops.groupnorm_nhwc(y, g, b, B, 112 * 112, 128, 8, ops.ACT_RELU, quad_mean=True))
print("groupnorm 112^2 x128 relu quad_mean: %.1f us (%.0f GB/s over 2 reads)" % (t, 2 * y.numel() * 4 / t * 1e-3), flush=True)
t = timed(lambda: ops.groupnorm_nhwc(y, g, b, B, 112 * 112, 128, 8, ops.ACT_SIGMOID))
print("groupnorm 112^2 x128 sigmoid: %.1f us (%.0f GB/s over 2 reads + 1 write)" % (t, 3 * y.numel() * 4 / t * 1e-3), flush=True)
z = torch.randn((B, 112, 112, 32), device=dev)
t = timed(lambda: ops.resample_nhwc(z, B, 112, 112, 32, ops.RS_UP_ALIGNED, 2))
print("upsample 112^2x32 -> 224^2: %.1f us (%.0f GB/s)" % (t, (z.numel() * 4 + 4 * z.numel() * 4) / t * 1e-3), flush=True)
