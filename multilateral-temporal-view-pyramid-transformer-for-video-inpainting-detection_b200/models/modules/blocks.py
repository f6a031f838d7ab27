"""ViT blocks of the temporal "global" encoder on libmumpy_b200 kernels.

Mirrors reference models/modules/blocks.py: FeedForward (:14-34), Attention (:37-74), Block (:77-92).
In Mumpy these run under vmap over the 49 spatial positions, i.e. on sequences of 3 temporal tokens
(multiTemporalViewEncoder.py:741), so attention is a 3x3 softmax per head and the work is four GEMMs.
"""
import torch.nn as nn

from ... import ops
from ._packing import PackedModule, as_operand, require_inference


class FeedForward(PackedModule):
    def __init__(self, dim, hidden_dim, dropout, out_dim=None):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden_dim)
        self.act = nn.GELU()
        if out_dim is None:
            out_dim = dim
        self.fc2 = nn.Linear(hidden_dim, out_dim)
        self.drop = nn.Dropout(dropout)

    @property
    def unwrapped(self):
        return self

    def fused(self, xn, residual=None):
        h = ops.linear(xn, self._gemm_weight("fc1", self.fc1.weight), self.fc1.bias, act=ops.ACT_GELU, out_dtype=ops.act_dtype())
        return ops.linear(h, self._gemm_weight("fc2", self.fc2.weight), self.fc2.bias, residual=residual)

    def forward(self, x):
        require_inference(self)
        return self.fused(as_operand(x.contiguous()))


class Attention(PackedModule):
    def __init__(self, dim, heads, dropout):
        super().__init__()
        self.heads = heads
        head_dim = dim // heads
        self.scale = head_dim ** -0.5
        self.attn = None
        self.qkv = nn.Linear(dim, dim * 3)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(dropout)

    @property
    def unwrapped(self):
        return self

    def fused(self, xn, Bn, N, C, residual=None, want_attn=False):
        qkv = ops.linear(xn, self._gemm_weight("qkv", self.qkv.weight), self.qkv.bias, out_dtype=ops.act_dtype())
        if want_attn:
            # the softmax map itself (blocks.py:66-68): a diagnostic output, computed by its own small kernel from the same q / k
            self.attn_map = ops.mha_short_probs(qkv, Bn, N, C, self.heads)
        ao = ops.mha_short(qkv, Bn, N, C, self.heads)
        return ops.linear(ao, self._gemm_weight("proj", self.proj.weight), self.proj.bias, residual=residual)

    def forward(self, x, mask=None):
        """Returns (x, attn) like the reference (blocks.py:54-74): attn (B, heads, N, N) fp32."""
        require_inference(self)
        B, N, C = x.shape
        y = self.fused(as_operand(x.contiguous()), B, N, C, want_attn=True)
        return y, self.attn_map


class Block(PackedModule):
    def __init__(self, dim, heads, mlp_dim, dropout, drop_path):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn = Attention(dim, heads, dropout)
        self.mlp = FeedForward(dim, mlp_dim, dropout)
        self.drop_path = nn.Identity()

    def forward(self, x, mask=None, return_attention=False):
        require_inference(self)
        B, N, C = x.shape
        x = x.contiguous()
        xn = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        if return_attention:                                     # blocks.py:88-89: the map of this block's attention, nothing else
            qkv = ops.linear(xn, self.attn._gemm_weight("qkv", self.attn.qkv.weight), self.attn.qkv.bias, out_dtype=ops.act_dtype())
            return ops.mha_short_probs(qkv, B, N, C, self.attn.heads)
        x = self.attn.fused(xn, B, N, C, residual=x)
        xn = ops.layernorm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return self.mlp.fused(xn, residual=x)
