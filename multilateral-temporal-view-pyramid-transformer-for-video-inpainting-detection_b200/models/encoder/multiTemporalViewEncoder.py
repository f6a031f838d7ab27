"""Multi-temporal-view Swin backbone on libmumpy_b200 kernels.

Mirrors the live classes of reference models/encoder/multiTemporalViewEncoder.py (same names, constructor and
forward signatures, state_dict keys): CVAModule (:127-139), CrossSwinBlock (:142-291), CrossThreeViewSwinBlock
(:294-350), OriginalThreeViewSwinBlock (:390-450), MultiViewBasicLayer (:489-538), CreateStages (:541-571),
CrossThreeViewTokenize (:574-618), CreateGlobalBlocks (:657-669), ThreeViewSwinTransformer (:672-746).
The reference's dead classes (CrossWindowAttention, CrossMultipleView*, OriginalTMultipleView*) are not provided.

Token tensors stay in canvas order (B, T*H*W, C) end to end; window partition / cyclic shift / window reverse
and the temporal-view regrouping are index maps inside the kernels, so no permute or roll copies are made.
"""
import torch
import torch.nn as nn

from ... import ops, streams
from ..modules._packing import PackedModule, as_operand, require_inference
from ..modules.blocks import Block
from ..modules.dct import FAF
from ..modules.deformableAttention import SwinDAttention
from ..modules.swinTransformer import (Mlp, SwinTransformerBlock, ThreeViewPatchMerging, WindowAttention, _shift_mask,
                                       to_2tuple)

# Pairing of query and key/value windows in the deformable cross-view attention: False reproduces the
# reference exactly (batch-global modulo, results depend on the batch a clip is in -- SURVEY finding 4a / A9);
# True applies the batch-1 map inside every clip so that results are independent of batching and sharding.
PER_CLIP_PAIRING = False


def set_per_clip_pairing(flag: bool):
    global PER_CLIP_PAIRING
    PER_CLIP_PAIRING = bool(flag)


class CVAModule(nn.Module):
    def __init__(self, dim1, num_heads, window_size=7, temporal_dims=[], qkv_bias=True, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., cur_stage=0):
        super().__init__()
        ws = window_size[0] if isinstance(window_size, (tuple, list)) else window_size
        # the reference leaves SwinDAttention.ws at its default 7 (:131); windows of another size need it set
        self.crossattn = SwinDAttention(dim1, num_heads, attn_drop, n_groups=3, ws=ws)
        self.drop_path = nn.Identity()

    def forward(self, x1, x2, mask=None, return_attention=False):
        """x1 (N1,P,C), x2 (N2,P,C) windows -> (x1 + y, attn), or attn alone with return_attention   (:134-139)."""
        y, attn = self.crossattn(x1, x2)
        if return_attention:
            return attn
        return ops.add(x1.contiguous().float(), y), attn


class CrossSwinBlock(PackedModule):
    def __init__(self, dim1, dim2, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 fused_window_process=False, last_view=False, temporal_dims=1, cur_stage=0):
        super().__init__()
        self.dim = dim1
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.last_view = last_view
        self.temporal_dims = temporal_dims
        self.cur_stage = cur_stage
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(self.dim)
        self.attn = WindowAttention(self.dim, window_size=to_2tuple(self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim1)
        self.mlp = Mlp(in_features=dim1, hidden_features=int(dim1 * mlp_ratio), act_layer=act_layer, drop=drop)
        self.pre = nn.Identity() if self.last_view else nn.Linear(dim2, dim1)
        if not self.last_view:
            nn.init.trunc_normal_(self.pre.weight, std=.02)
            nn.init.zeros_(self.pre.bias)
        self.cva = nn.Identity() if self.last_view else CVAModule(dim1, temporal_dims=self.temporal_dims,
                                                                  window_size=to_2tuple(self.window_size),
                                                                  num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                                                  attn_drop=attn_drop, drop=drop, drop_path=drop_path,
                                                                  cur_stage=cur_stage)
        attn_mask = None
        if self.shift_size > 0:
            H, W = self.input_resolution
            attn_mask = _shift_mask(H, W, self.temporal_dims, self.window_size, self.shift_size)
        self.register_buffer("attn_mask", attn_mask)
        self.fused_window_process = fused_window_process

    def attn_phase(self, x1, need_out=True, need_out_fp32=False):
        """LN1 -> W-MSA -> proj on the canvas x1 (B,L1,C1) fp32 -> (h, out_fp32 or None, out_op): h = x1 + out is the
        shortcut sum (:276), out the un-summed W-MSA branch (:275) and out_op its GEMM-operand form, which is all the
        next view's `pre` needs.  In bf16 mode the projection writes h (fp32) and out_op (bf16) from the same
        accumulators (mumpy_linear_dual)."""
        H, W = self.input_resolution
        B, L1, C1 = x1.shape
        TH1 = L1 // W
        x1 = x1.contiguous()
        ao = self.attn.canvas_attention(x1, B, TH1, W, self.shift_size, self.attn_mask, norm=self.norm1)
        if ops.tensor_cores() and not need_out_fp32:
            if need_out:
                h, out_op = ops.linear_dual(ao, self.attn._gemm_weight("proj", self.attn.proj.weight), self.attn.proj.bias, x1)
            else:
                h, out_op = self.attn.project(ao, residual=x1), None
            return h, None, out_op
        out = self.attn.project(ao)                                   # `out` (:275)
        h = ops.add(x1, out)                                          # shortcut + attn (:276)
        return h, out, (as_operand(out) if need_out else None)

    def tail_phase(self, h, x2_op):
        """Deformable cross-view attention against the other view's branch output (skipped for the last view), then the MLP."""
        H, W = self.input_resolution
        B, L1, C1 = h.shape
        TH1 = L1 // W
        if not self.last_view:
            TH2 = x2_op.shape[1] // W
            # `pre` is per token, so it commutes with window_partition: apply it on the canvas (:282-283); its result is
            # only ever bilinearly sampled, so it is kept in operand precision
            x2p = ops.linear(x2_op, self._gemm_weight("pre", self.pre.weight), self.pre.bias, out_dtype=ops.act_dtype())
            y = self.cva.crossattn.canvas_forward(h, x2p, B, TH1, TH2, W, PER_CLIP_PAIRING)
            # h + (window_partition(h) + raw_reshape(y)) added in window-major order, no window_reverse (:138,284-286)
            h = ops.cva_residual(h, y, B, TH1, W, C1, self.window_size)
        return self.mlp.fused(h, residual=h, norm=self.norm2)

    def forward(self, x1, x2):
        """x1 (B,L1,C1), x2 (B,L2,C2) canvases -> (x1', out) with out = the un-summed W-MSA branch (:228-291)."""
        require_inference(self)
        x2_op = None if self.last_view else as_operand(x2.contiguous())
        h, out, _ = self.attn_phase(x1, need_out=False, need_out_fp32=True)
        return self.tail_phase(h, x2_op), out


class CrossThreeViewSwinBlock(nn.Module):
    def __init__(self, view_configs, input_resolution, cur_stage, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, fused_window_process=False):
        super().__init__()
        common = dict(shift_size=0, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                      attn_drop=attn_drop, drop_path=drop_path, act_layer=nn.GELU, fused_window_process=fused_window_process,
                      cur_stage=cur_stage)
        hs = [view_configs[i]["hidden_size"][cur_stage] for i in range(3)]
        nh = [view_configs[i]["num_heads"][cur_stage] for i in range(3)]
        ws = [view_configs[i]["window_size"] for i in range(3)]
        self.block1 = CrossSwinBlock(hs[0], hs[1], input_resolution[0], nh[0], window_size=ws[0], norm_layer=norm_layer,
                                     temporal_dims=view_configs[0]["temporal_ratio"], **common)
        self.block2 = CrossSwinBlock(hs[1], hs[2], input_resolution[1], nh[1], window_size=ws[1], norm_layer=nn.LayerNorm,
                                     temporal_dims=view_configs[0]["temporal_ratio"], **common)
        self.block3 = CrossSwinBlock(hs[2], hs[2], input_resolution[2], nh[2], window_size=ws[2], norm_layer=nn.LayerNorm,
                                     last_view=True, temporal_dims=3, **common)

    def forward(self, x):
        # order view3 -> view2 (+CVA from view3) -> view1 (+CVA from view2)   (:345-350)
        for blk in (self.block1, self.block2, self.block3):
            require_inference(blk)
        # the three attention phases are independent; only the deformable cross-view step of view v waits for the
        # attention branch of view v+1, so each view runs on its own lane with two event hand-overs
        with streams.region(x[0].device) as reg:
            with reg.lane(2):
                h3, _, op3 = self.block3.attn_phase(x[2])
                reg.publish((id(self), 3), op3)
                x[2] = self.block3.tail_phase(h3, None)
            with reg.lane(1):
                h2, _, op2 = self.block2.attn_phase(x[1])
                reg.publish((id(self), 2), op2)
                reg.need((id(self), 3))
                x[1] = self.block2.tail_phase(h2, op3)
            with reg.lane(0):
                h1, _, _ = self.block1.attn_phase(x[0], need_out=False)
                reg.need((id(self), 2))
                x[0] = self.block1.tail_phase(h1, op2)
        return x


class OriginalThreeViewSwinBlock(nn.Module):
    def __init__(self, view_configs, input_resolution, cur_stage, cur_lyr, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, fused_window_process=False):
        super().__init__()
        for i in range(3):
            if cur_lyr < view_configs[i]["depths"][cur_stage]:
                blk = SwinTransformerBlock(dim=view_configs[i]["hidden_size"][cur_stage], input_resolution=input_resolution[i],
                                           num_heads=view_configs[i]["num_heads"][cur_stage],
                                           window_size=view_configs[i]["window_size"],
                                           shift_size=0 if (cur_lyr % 2 == 0) else view_configs[0]["window_size"] // 2,
                                           mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                                           attn_drop=attn_drop, drop_path=drop_path, norm_layer=norm_layer,
                                           fused_window_process=fused_window_process,
                                           temporal_dim=view_configs[i]["temporal_dim"])
            else:
                blk = nn.Identity()                                   # view past its own depth (:415)
            setattr(self, "block%d" % (i + 1), blk)

    def forward(self, x):
        with streams.region(x[2].device) as reg:          # the three views are independent: one lane each
            with reg.lane(2):
                x[2] = self.block3(x[2])
            with reg.lane(1):
                x[1] = self.block2(x[1])
            with reg.lane(0):
                x[0] = self.block1(x[0])
        return x


class MultiViewBasicLayer(nn.Module):
    def __init__(self, view_configs, cur_stage, depth, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., norm_layer=nn.LayerNorm, downsample=None, fused_window_process=False):
        super().__init__()
        res = [view_configs[k]["input_resolution"][cur_stage] for k in range(3)]
        blocks = []
        for i in range(depth):
            dp = drop_path[i] if isinstance(drop_path, list) else drop_path
            kw = dict(cur_stage=cur_stage, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                      attn_drop=attn_drop, drop_path=dp, norm_layer=norm_layer, fused_window_process=fused_window_process)
            blocks.append(CrossThreeViewSwinBlock(view_configs, input_resolution=res, **kw) if i == 0 else
                          OriginalThreeViewSwinBlock(view_configs, input_resolution=res, cur_lyr=i, **kw))
        self.blocks = nn.ModuleList(blocks)
        self.downsample = downsample(view_configs, cur_stage) if downsample is not None else None

    def forward(self, x):
        out = []
        with streams.region(x[0].device) as reg:
            for blk in self.blocks:
                x = blk(x)
                out = x.copy()
            for v in range(len(out)):            # the stage features are consumed across lanes (decoder pyramid): publish them
                with reg.lane(v):
                    reg.publish_tensor(out[v])
            if self.downsample is not None:
                x = self.downsample(x)
        return x, out


class CreateStages(nn.Module):
    def __init__(self, view_configs, depths=[2, 2, 18, 2], mlp_ratio=4., qkv_bias=True, qk_scale=None, stages=4,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, fused_window_process=False):
        super().__init__()
        self.layers = nn.ModuleList()
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        for i in range(stages):
            self.layers.append(MultiViewBasicLayer(view_configs=view_configs, cur_stage=i, depth=depths[i], mlp_ratio=mlp_ratio,
                                                   qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                                                   drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                                                   downsample=ThreeViewPatchMerging if (i < stages - 1) else None,
                                                   fused_window_process=fused_window_process))

    def forward(self, x):
        out = []
        with streams.region(x[0].device):
            for lyr in self.layers:
                x, out_stage = lyr(x)
                out.append(out_stage)
        return x, out


TENSOR_CORE_TOKENIZER = True      # 16-bit modes: tokenizer convolution as split-operand tcgen05 GEMMs (False: the fp32 FMA kernel)


class CrossThreeViewTokenize(PackedModule):
    def __init__(self, view_configs):
        super().__init__()
        for i in range(3):
            ps = view_configs[i]["patches"].size if hasattr(view_configs[i]["patches"], "size") else view_configs[i]["patches"]["size"]
            k = (ps[-1], ps[0], ps[1])
            setattr(self, "project%d" % (i + 1), nn.Conv3d(3, view_configs[i]["hidden_size"][0], kernel_size=k, stride=k, padding=0))
        for i in range(3):
            setattr(self, "norm%d" % (i + 1), nn.LayerNorm(view_configs[i]["hidden_size"][0]))

    def forward(self, x):
        """x (B,T,3,S,S) -> [ (B,T_v,HW,C_v) ] for the three views (:605-618)."""
        require_inference(self)
        x = x.contiguous().float()
        B, T, _, S, _ = x.shape
        out = []
        with streams.region(x.device) as reg:
            for i in range(3):
                proj = getattr(self, "project%d" % (i + 1))
                norm = getattr(self, "norm%d" % (i + 1))
                kt = proj.kernel_size[0]
                if proj.kernel_size[1:] != (4, 4):
                    raise NotImplementedError("tokenizer kernel supports 4x4 spatial patches")
                C = proj.out_channels
                with reg.lane(i):
                    if ops.tensor_cores() and TENSOR_CORE_TOKENIZER:
                        # patches [hi | hi | lo] x weight rows [w_hi | w_lo | w_hi] on the tensor cores (fp32 accuracy: the tokens seed
                        # the fp32 residual stream), + bias, then LayerNorm in fp32
                        def split_w(p=proj, c=C):
                            w = p.weight.detach().reshape(c, -1).contiguous()
                            hi = w.to(ops.act_dtype())
                            lo = (w - hi.float()).to(ops.act_dtype())
                            return torch.cat([hi, lo, hi], dim=1).contiguous()
                        w3 = self._packed("w3_%d" % i, [proj.weight], split_w)
                        tok = ops.linear(ops.patchify16(x, kt), w3, proj.bias)
                        tok = ops.layernorm(tok, norm.weight, norm.bias, norm.eps, out_dtype=torch.float32)
                    else:
                        w_kc = self._packed("w%d" % i, [proj.weight], lambda p=proj, c=C: p.weight.detach().reshape(c, -1).t().contiguous())
                        tok = ops.tokenize(x, w_kc, proj.bias, norm.weight, norm.bias, kt, norm.eps)
                out.append(tok.view(B, T // kt, (S // 4) ** 2, C))
        return out


class CreateGlobalBlocks(nn.Module):
    def __init__(self, global_encoder_config, dpr, dropout_rate):
        super().__init__()
        self.blocks = nn.ModuleList([Block(global_encoder_config["hidden_size"], global_encoder_config["num_heads"],
                                           global_encoder_config["mlp_dim"], dropout_rate, dpr[i])
                                     for i in range(global_encoder_config["num_layers"])])

    def forward(self, x):
        for block in self.blocks:
            x = block(x)
        return x


class ThreeViewSwinTransformer(PackedModule):
    def __init__(self, view_configs, input_token_temporal_dims, global_encoder_config, depths=[2, 2, 18, 2], mlp_ratio=4.,
                 qkv_bias=True, qk_scale=None, stages=4, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True, use_checkpoint=False, fused_window_process=False):
        super().__init__()
        size = view_configs[0]["input_resolution"][0][0] * 4
        self.faf = FAF(size)
        self.tokenize = CrossThreeViewTokenize(view_configs)
        self.input_token_temporal_dims = input_token_temporal_dims
        self.layers = CreateStages(view_configs, depths=depths, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                   stages=stages, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                                   drop_path_rate=drop_path_rate, norm_layer=norm_layer, ape=ape, patch_norm=patch_norm,
                                   use_checkpoint=use_checkpoint, fused_window_process=fused_window_process)
        self.pos_drop = nn.Dropout(p=drop_rate)
        merged = sum(view_configs[i]["hidden_size"][-1] for i in range(3))
        self.globalembedding = nn.Linear(merged, global_encoder_config["hidden_size"])
        self.global_dpr = [x.item() for x in torch.linspace(0, drop_path_rate, global_encoder_config["num_layers"])]
        self.globalblocks = CreateGlobalBlocks(global_encoder_config, self.global_dpr, drop_rate)

    def forward(self, x):
        """x (B,3,3,S,S) -> (x (B,n,3*768), out_x [4][3] of (B,1,L,C), ffinfo (B,9,S,S))   (:732-746)."""
        require_inference(self)
        x = x.contiguous().float()
        B = x.shape[0]
        with streams.region(x.device) as reg:
            with reg.lane(3):
                ffinfo = self.faf.frame(x, 1)                              # == faf(x)[:, 1]  
                reg.publish_tensor(ffinfo)
            toks = self.tokenize(x)
            # align_temporal_dimension_across_views + vmap over the size-1 dim == flatten (t, n) (:701-708,737)
            xs = [t.reshape(B, -1, t.shape[-1]) for t in toks]
            xs, outs = self.layers(xs)
            out_x = [[t.unsqueeze(1) for t in stage] for stage in outs]
            # the temporal global encoder runs on the caller's stream as soon as the three view lanes have delivered stage 3;
            # inside an outer region (mumpy_b200.forward) the frequency lane and the decoder's branches keep running beside it
            reg.wait_lanes([0, 1, 2])
            g = self._global_part(xs, B)
        return g, out_x, ffinfo

    def _global_part(self, xs, B):
        x = xs[0]
        # merge_views_along_channel_axis + globalembedding (:710-718,740); rows ordered (b, n, t)
        T = max(self.input_token_temporal_dims)
        n = xs[0].shape[1]
        widths = [t.shape[-1] for t in xs]
        merged = torch.empty((B * n * T, sum(widths)), dtype=ops.act_dtype(), device=x.device)
        col = 0
        for v, t in enumerate(xs):
            if self.input_token_temporal_dims[v] == 1:
                ops.gather_rows(t, widths[v], merged, merged.shape[1], col, B, n * T, n, div=T, mul_hi=1, mul_lo=0)
            else:
                ops.gather_rows(t, widths[v], merged, merged.shape[1], col, B, n * T, n * T, div=T, mul_hi=1, mul_lo=n)
            col += widths[v]
        g = ops.linear(merged, self._gemm_weight("globalembedding", self.globalembedding.weight), self.globalembedding.bias)
        # vmap(globalblocks, in_dims=2) == blocks on (B*n, T, 768)  (:741)
        g = self.globalblocks(g.view(B * n, T, -1))
        return g.view(B, n, -1)                                          # == cat over t on the channel axis (:745)
