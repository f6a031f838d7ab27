"""Window-level deformable cross-view attention on libmumpy_b200 kernels.

Mirrors reference models/modules/deformableAttention.py: LayerNormProxy (:11-19) and SwinDAttention (:218-405)
(the unused, broken DAttention :25-212 is not provided).  The layout quirks the weights were trained with are
reproduced exactly (SURVEY finding 4): query window (r*i+t) mod N1 pairs with kv window r*i+t, the sum over the
temporal ratio, and the raw (C,7,7)->(49,C) reinterpretation of the output (:403).
"""
import torch
import torch.nn as nn

from ... import ops
from ._packing import PackedModule, as_operand, require_inference


class LayerNormProxy(nn.Module):
    """Parameter container for conv_offset.1 (LayerNorm over the group channels); evaluated inside cva_offsets."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        raise RuntimeError("LayerNormProxy is fused into the offset kernel and is not called on its own")


class SwinDAttention(PackedModule):
    def __init__(self, dim1, n_heads, attn_drop, n_groups, ws=7, stride=1, offset_range_factor=2, no_off=False,
                 height_scale=[1, 1], dwc_pe=False, use_pe=False, fixed_pe=False):
        super().__init__()
        if stride != 1 or offset_range_factor != 2 or no_off or use_pe or dwc_pe or fixed_pe:
            raise NotImplementedError("only the configuration Mumpy constructs (CVAModule) is implemented")
        self.n_head_channels = dim1 // n_heads
        self.scale = self.n_head_channels ** -0.5
        self.n_heads = n_heads
        self.ws = ws
        self.nc = self.n_head_channels * n_heads
        self.n_groups = n_groups
        self.n_group_channels = self.nc // self.n_groups
        self.n_group_heads = self.n_heads // self.n_groups
        self.no_off = no_off
        self.offset_range_factor = offset_range_factor
        self.use_pe = use_pe
        kk = 5
        cg = self.n_group_channels
        self.conv_offset = nn.Sequential(
            nn.Conv2d(cg, cg, kk, stride, kk // 2, groups=cg),
            LayerNormProxy(cg),
            nn.GELU(),
            nn.Conv2d(cg, 2, 1, 1, 0, bias=False))
        self.proj_q = nn.Conv2d(self.nc, self.nc, kernel_size=1, stride=1, padding=0)
        self.proj_k = nn.Conv2d(self.nc, self.nc, kernel_size=1, stride=1, padding=0)
        self.proj_v = nn.Conv2d(self.nc, self.nc, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(self.nc, self.nc, kernel_size=1, stride=1, padding=0)
        self.proj_drop = nn.Dropout(attn_drop, inplace=False)
        self.attn_drop = nn.Dropout(attn_drop, inplace=False)
        self.rpe_table = None
        for p in (self.proj_q, self.proj_k, self.proj_v):          # deformableAttention.py:302-309
            nn.init.trunc_normal_(p.weight)
            nn.init.zeros_(p.bias)
        nn.init.zeros_(self.proj_out.weight)
        nn.init.zeros_(self.proj_out.bias)

    # ---- packed operands -------------------------------------------------------------------------------------
    def _kv_weight(self):
        def make():
            w = torch.cat([self.proj_k.weight.detach().reshape(self.nc, self.nc),
                           self.proj_v.weight.detach().reshape(self.nc, self.nc)], 0).contiguous()
            return ops.cast16(w) if ops.tensor_cores() else w
        return self._packed("kv_w", [self.proj_k.weight, self.proj_v.weight], make)

    def _kv_bias(self):
        return self._packed("kv_b", [self.proj_k.bias, self.proj_v.bias],
                            lambda: torch.cat([self.proj_k.bias.detach(), self.proj_v.bias.detach()]).contiguous())

    def _offset_params(self):
        cg = self.n_group_channels
        return self._packed("off", [self.conv_offset[0].weight, self.conv_offset[3].weight],
                            lambda: (self.conv_offset[0].weight.detach().reshape(cg, 25).contiguous(),
                                     self.conv_offset[3].weight.detach().reshape(2, cg).contiguous()))

    # ---- canvas form used by CrossSwinBlock ------------------------------------------------------------------
    def canvas_forward(self, h, x2p, B, TH1, TH2, W, per_clip_pairing=False, want_attn=False):
        """h (B,TH1*W,C) fp32 query canvas, x2p (B,TH2*W,C) fp32 key/value canvas (after `pre`).
        Returns proj_out's token-major result y (N1*P, C) fp32; the raw reshape of :403 is applied by
        ops.cva_residual, which consumes y."""
        C, ws = self.nc, self.ws
        q = ops.linear(as_operand(h), self._gemm_weight("q", self.proj_q.weight, (C, C)), self.proj_q.bias)
        dw_w, pw = self._offset_params()
        ln = self.conv_offset[1].norm
        pix = ops.cva_offsets(q, dw_w, self.conv_offset[0].bias, ln.weight, ln.bias, pw, B, TH1, W, C, self.n_groups, ws)
        samp = ops.cva_sample(x2p, pix, B, TH1, TH2, W, C, self.n_groups, ws, per_clip_pairing, ops.act_dtype())
        kv = ops.linear(samp, self._kv_weight(), self._kv_bias(), out_dtype=ops.act_dtype())
        o = ops.cva_attention(q, kv, B, TH1, TH2, W, C, self.n_heads, ws, per_clip_pairing)
        y = ops.linear(o, self._gemm_weight("out", self.proj_out.weight, (C, C)), self.proj_out.bias)
        if want_attn:          # the reference's second result (:399,405): (N1, r * heads, P, P), from the same q / sampled k
            return y, ops.cva_attention_probs(q, kv, B, TH1, TH2, W, C, self.n_heads, ws, per_clip_pairing)
        return y

    def forward(self, x1, x2, return_attention=False):
        """x1 (N1, ws*ws, C) query windows, x2 (N2, ws*ws, C) key/value windows -> (x (N1, ws*ws, C), attn (N1, r*heads, P, P)) like the
        reference (:324-405; `return_attention` is accepted and ignored there too)."""
        require_inference(self)
        N1, P, C = x1.shape
        N2 = x2.shape[0]
        ws = self.ws
        x1 = x1.contiguous().float()
        x2 = x2.contiguous().float()
        # a stack of n windows is an (n*ws) x ws canvas whose windows are exactly the inputs, in order
        y, attn = self.canvas_forward(x1.view(1, N1 * P, C), x2.view(1, N2 * P, C), 1, N1 * ws, N2 * ws, ws, want_attn=True)
        zero = torch.zeros_like(x1).view(1, N1 * P, C)
        return ops.cva_residual(zero, y, 1, N1 * ws, ws, C, ws).view(N1, P, C), attn
