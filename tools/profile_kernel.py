"""Development aid: launch one kernel family in isolation for `ncu --set full -k regex:<name>`.
    python tools/profile_kernel.py attn|gemm_fc1|gemm_fc2|gemm_s0|conv|ln [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mumpy_b200 import ops  # noqa: E402
from oracle import mumpy_oracle as orc  # noqa: E402

what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
torch.manual_seed(0)
reps = 3
if what == "attn":
    C, heads, TH, W = 128, 4, 168, 56
    qkv = torch.randn((B, TH * W, 3 * C), device=dev).bfloat16()
    bias = torch.randn((heads, 49, 49), device=dev)
    mask = orc.shifted_window_mask(TH, W, 7, 3).to(dev)
    table = torch.randn((169, heads), device=dev)
    for _ in range(reps):
        ops.window_attention(qkv, bias, mask, B, TH, W, C, heads, 7, 3, rel_table=table, standard_mask=True)
elif what.startswith("gemm"):
    shapes = {"gemm_fc1": (B * 588, 2048, 512, ops.ACT_GELU, torch.bfloat16), "gemm_fc2": (B * 588, 512, 2048, ops.ACT_NONE, torch.float32),
              "gemm_s0": (B * 9408, 512, 128, ops.ACT_GELU, torch.bfloat16), "gemm_qkv": (B * 588, 1536, 512, ops.ACT_NONE, torch.bfloat16),
              "gemm_small": (B * 196, 384, 384, ops.ACT_NONE, torch.float32), "gemm_small_fc1": (B * 196, 1536, 384, ops.ACT_GELU, torch.bfloat16)}
    M, N, K, act, odt = shapes[what]
    a = torch.randn((M, K), device=dev).bfloat16()
    w = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
    bias = torch.zeros(N, device=dev)
    res = torch.randn((M, N), device=dev) if odt == torch.float32 else None
    for _ in range(reps):
        ops.linear(a, w, bias, res, act=act, out_dtype=odt)
elif what == "ln":
    x = torch.randn((B * 9408, 128), device=dev)
    g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    for _ in range(reps):
        ops.layernorm(x, g, b, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("done")
