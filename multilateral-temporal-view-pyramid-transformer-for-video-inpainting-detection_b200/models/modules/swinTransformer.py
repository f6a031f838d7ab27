"""Swin modules of the Mumpy encoder on libmumpy_b200 kernels.

Mirrors the live classes of reference models/modules/swinTransformer.py (same names, constructor and forward
signatures, state_dict keys): Mlp (:35-51), window_partition/window_reverse (:54-83), WindowAttention (:86-166),
SwinTransformerBlock (:185-307), PatchMerging (:328-367), ThreeViewPatchMerging (:637-657).
Inference only (DropPath / dropout are identities in eval()); arithmetic is done by the CUDA library.
"""
import torch
import torch.nn as nn

from ... import ops, streams
from ._packing import PackedModule, as_operand, require_inference


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def window_partition(x, window_size):
    """(B,H,W,C) -> (num_windows*B, ws, ws, C).  Layout helper for API parity; the kernels never call it."""
    B, H, W, C = x.shape
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size, H, W):
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


class Mlp(PackedModule):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("libmumpy_b200 Mlp supports nn.GELU only (the reference never uses another)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def fused(self, xn, residual=None, norm=None):
        """xn: normalised input already in operand precision -- or, with `norm` (the block's LayerNorm), the raw fp32 stream, which
        is then normalised inside the fc1 kernel where the width fits (mumpy_ln_linear); returns fc2(gelu(fc1(.))) (+ residual), fp32."""
        w1 = self._gemm_weight("fc1", self.fc1.weight)
        if (norm is not None and residual is xn and ops.mlp_fused_fits(w1.shape[1]) and w1.shape[0] == 4 * w1.shape[1]
                and self.fc1.bias is not None and self.fc2.bias is not None and tuple(self.fc2.weight.shape) == (w1.shape[1], w1.shape[0])):
            # narrow (HBM-bound) stages: the whole `x + Mlp(LN(x))` in one kernel, hidden activations never leave the SM
            return ops.mlp_fused(xn, norm.weight, norm.bias, norm.eps, w1, self.fc1.bias, self._gemm_weight("fc2", self.fc2.weight), self.fc2.bias)
        if norm is not None and ops.ln_linear_fits(w1.shape[0], w1.shape[1]):
            h = ops.ln_linear(xn, norm.weight, norm.bias, norm.eps, w1, self.fc1.bias, act=ops.ACT_GELU)
        else:
            if norm is not None:
                xn = ops.layernorm(xn, norm.weight, norm.bias, norm.eps)
            h = ops.linear(xn, w1, self.fc1.bias, act=ops.ACT_GELU, out_dtype=ops.act_dtype())
        return ops.linear(h, self._gemm_weight("fc2", self.fc2.weight), self.fc2.bias, residual=residual)

    def forward(self, x):
        require_inference(self)
        return self.fused(as_operand(x.contiguous()))


class WindowAttention(PackedModule):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if qk_scale is not None and abs(qk_scale - head_dim ** -0.5) > 1e-12:
            raise NotImplementedError("custom qk_scale is not supported by the fused kernel")
        self.scale = head_dim ** -0.5
        ws0, ws1 = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws0 - 1) * (2 * ws1 - 1), num_heads))
        coords = torch.stack(torch.meshgrid([torch.arange(ws0), torch.arange(ws1)], indexing="ij"))
        cf = torch.flatten(coords, 1)
        rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws0 - 1
        rel[:, :, 1] += ws1 - 1
        rel[:, :, 0] *= 2 * ws1 - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)

    def _bias(self):
        """(nH, N, N) fp32 = table[index] -- input independent, gathered once per weight load (:148-150)."""
        def make():
            n = self.window_size[0] * self.window_size[1]
            t = self.relative_position_bias_table.detach()[self.relative_position_index.view(-1)]
            return t.view(n, n, -1).permute(2, 0, 1).contiguous().float()
        return self._packed("bias", [self.relative_position_bias_table, self.relative_position_index], make)

    def _table(self):
        return self._packed("table", [self.relative_position_bias_table],
                            lambda: self.relative_position_bias_table.detach().float().contiguous())

    def canvas_attention(self, xn, B, TH, W, shift, mask, standard_mask=False, norm=None):
        """xn (B, TH*W, C) normalised canvas in operand precision -> attention output before `proj`.  With `norm` (the block's
        norm1) xn is the raw fp32 canvas and the LayerNorm runs inside the qkv kernel where the width fits (mumpy_ln_linear).
        standard_mask: `mask` is exactly the Swin shift mask for (TH, W, ws, shift) (lets the kernel recompute it)."""
        C = self.dim
        wq = self._gemm_weight("qkv", self.qkv.weight)
        if norm is not None and ops.ln_linear_fits(wq.shape[0], wq.shape[1]):
            qkv = ops.ln_linear(xn, norm.weight, norm.bias, norm.eps, wq, self.qkv.bias)
        else:
            if norm is not None:
                xn = ops.layernorm(xn, norm.weight, norm.bias, norm.eps)
            qkv = ops.linear(xn, wq, self.qkv.bias, out_dtype=ops.act_dtype())
        return ops.window_attention(qkv, self._bias(), mask, B, TH, W, C, self.num_heads, self.window_size[0], shift,
                                    rel_table=self._table(), standard_mask=standard_mask)

    def project(self, ao, residual=None):
        return ops.linear(ao, self._gemm_weight("proj", self.proj.weight), self.proj.bias, residual=residual)

    def forward(self, x, mask=None):
        """x (num_windows*B, N, C) already partitioned; mask (nW, N, N) or None  (:134-166)."""
        require_inference(self)
        B_, N, C = x.shape
        ws = self.window_size[0]
        nW = 1 if mask is None else mask.shape[0]
        # a stack of nW windows is a (nW*ws) x ws canvas whose windows are exactly the inputs, in order
        ao = self.canvas_attention(as_operand(x.contiguous()), B_ // nW, nW * ws, ws, 0,
                                   None if mask is None else mask.contiguous().float())
        return self.project(ao).view(B_, N, C)


def _shift_mask(H, W, temporal_dim, window_size, shift_size):
    """swinTransformer.py:233-252 on the stacked (temporal_dim*H) x W canvas."""
    img_mask = torch.zeros((1, H * temporal_dim, W, 1))
    slices = (slice(0, -window_size), slice(-window_size, -shift_size), slice(-shift_size, None))
    cnt = 0
    for h in slices:
        for w in slices:
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img_mask, window_size).view(-1, window_size * window_size)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


class SwinTransformerBlock(PackedModule):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 temporal_dim=1, fused_window_process=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.temporal_dim = temporal_dim
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()      # stochastic depth is the identity at inference
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        attn_mask = None
        if self.shift_size > 0:
            H, W = self.input_resolution
            attn_mask = _shift_mask(H, W, self.temporal_dim, self.window_size, self.shift_size)
        self.register_buffer("attn_mask", attn_mask)
        self.fused_window_process = fused_window_process

    def _mask_is_standard(self, TH, W):
        """True when the attn_mask buffer equals the mask the constructor builds for this canvas (checked once per load)."""
        if self.attn_mask is None:
            return False

        def check():
            ref = _shift_mask(TH, W, 1, self.window_size, self.shift_size).to(self.attn_mask.device)
            return bool(ref.shape == self.attn_mask.shape and torch.equal(ref, self.attn_mask))
        return self._packed("std_mask:%d:%d" % (TH, W), [self.attn_mask], check)

    def forward(self, x):
        require_inference(self)
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L % (H * W) == 0, "input feature has wrong size"
        TH = L // W
        x = x.contiguous()
        ao = self.attn.canvas_attention(x, B, TH, W, self.shift_size, self.attn_mask, self._mask_is_standard(TH, W), norm=self.norm1)
        x = self.attn.project(ao, residual=x)                    # shortcut + W-MSA  (:302)
        return self.mlp.fused(x, residual=x, norm=self.norm2)    # x + Mlp(LN2(x))   (:305)


class PatchMerging(PackedModule):
    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):
        require_inference(self)
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        xn = ops.patch_merge_norm(x.contiguous(), self.norm.weight, self.norm.bias, B, H, W, C, self.norm.eps)
        return ops.linear(xn, self._gemm_weight("reduction", self.reduction.weight))


class ThreeViewPatchMerging(nn.Module):
    def __init__(self, view_configs, cur_stage):
        super().__init__()
        for i in range(3):
            r = view_configs[i]["input_resolution"][cur_stage][0]
            setattr(self, "downsample%d" % (i + 1),
                    PatchMerging((view_configs[i]["temporal_dim"] * r, r), view_configs[i]["hidden_size"][cur_stage]))

    def forward(self, x):
        with streams.region(x[0].device) as reg:
            with reg.lane(0):
                x[0] = self.downsample1(x[0])
            with reg.lane(1):
                x[1] = self.downsample2(x[1])
            with reg.lane(2):
                x[2] = self.downsample3(x[2])
        return x
