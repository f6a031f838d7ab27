"""Multi-pyramid decoder on libmumpy_b200 kernels.

Mirrors reference models/decoder/decoder.py: SEB (:6-14), _GlobalConvModule (:17-39), Decoder (:41-225) with the same
constructor defaults, forward signature and state_dict keys (BaselineDecoder is an ablation and is not provided).
All maps are NHWC inside; convolutions are implicit GEMMs (exact fp32 in 'fp32' mode, im2col + tcgen05 GEMM in
'bf16' mode), GroupNorm statistics are fp32, and resampling / gating / skip additions are fused where they meet.
"""
import torch
import torch.nn as nn

from ... import ops, streams
from ..modules._packing import PackedModule, require_inference


def _conv(owner, name, conv, x, B, H, W, ld_in=None, out16=False, residual=None):
    """Stride-1 nn.Conv2d on an NHWC map x (B,H,W,Cin[ld_in]) -> (B,H,W,Cout) fp32 (+ residual, fused into the epilogue).
    out16: in the tensor-core modes return the result in operand precision (it only feeds another convolution)."""
    Cout, Cin, kh, kw = conv.weight.shape
    ph, pw = conv.padding
    K = kh * kw * Cin
    if Cout == 1 and Cin % 4 == 0 and ld_in in (None, Cin):
        w = owner._packed("hwi:" + name, [conv.weight], lambda: conv.weight.detach().permute(0, 2, 3, 1).contiguous())
        return ops.conv2d_nhwc_cout1(x, w, conv.bias, B, H, W, Cin, kh, kw, ph, pw)
    if ops.tensor_cores() and Cout % 8 == 0 and (ld_in or Cin) % 8 == 0:
        # implicit GEMM on the tensor cores: im2col-mode TMA feeds tcgen05 directly (no im2col buffer).  Cin itself may be
        # anything (the 9-channel frequency map): the tensor map's channel extent is Cin, TMA zero-fills the rest of the
        # 64-channel block, only the pixel stride has to be 16-byte aligned
        cb = (Cin + 63) // 64

        def make_tc():
            w = conv.weight.detach().permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)
            wp = w.new_zeros(Cout, kh * kw, cb * 64)
            wp[:, :, :Cin] = w
            return ops.cast16(wp.reshape(Cout, -1).contiguous())
        wq = owner._packed("tcconv:" + name, [conv.weight], make_tc)
        xb = x if x.dtype == ops.act_dtype() else ops.cast16(x)
        return ops.conv2d_nhwc_bf16(xb, wq, conv.bias, B, H, W, Cin, Cout, kh, kw, ph, pw, ld_in,
                                    out_dtype=ops.act_dtype() if out16 else torch.float32, residual=residual)
    if ops.tensor_cores() and Cout % 8 == 0:
        Kpad = (K + 7) // 8 * 8

        def make():
            w = conv.weight.detach().permute(0, 2, 3, 1).reshape(Cout, K)
            if Kpad != K:
                w = torch.cat([w, w.new_zeros(Cout, Kpad - K)], 1)
            return ops.cast16(w.contiguous())
        wq = owner._packed("tc:" + name, [conv.weight], make)
        cols = ops.im2col_nhwc(x, B, H, W, Cin, kh, kw, ph, pw, Kpad, ld_in)
        return ops.linear(cols, wq, conv.bias, residual=None if residual is None else residual.view(B * H * W, Cout)).view(B, H, W, Cout)
    w = owner._packed("ohwi:" + name, [conv.weight], lambda: conv.weight.detach().permute(0, 2, 3, 1).contiguous())
    y = ops.conv2d_nhwc(x, w, conv.bias, B, H, W, Cin, Cout, kh, kw, ph, pw, ld_in)
    return y if residual is None else ops.add(y, residual)


def _conv_in_dtype():
    """dtype of a map whose only consumer is a convolution: the GEMM operand type in the tensor-core modes (the producer writes
    it directly, no separate cast pass), fp32 otherwise."""
    return ops.act_dtype() if ops.tensor_cores() else torch.float32


class SEB(PackedModule):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear")

    def nhwc(self, x1, x2, B, H2, W2, out_dtype=torch.float32):
        """x1 (B,2H2,2W2,Co), x2 (B,H2,W2,Ci) NHWC -> x1 * up2(conv(x2)), half-pixel bilinear (:12-14)."""
        c = _conv(self, "conv", self.conv, x2, B, H2, W2)
        return ops.resample_nhwc(c, B, H2, W2, c.shape[-1], ops.RS_UP_HALFPIX, 2, mul=x1, out_dtype=out_dtype)

    def forward(self, x):
        require_inference(self)
        x1, x2 = x
        B, _, H2, W2 = x2.shape
        out = self.nhwc(ops.nchw_to_nhwc(x1.contiguous()), ops.nchw_to_nhwc(x2.contiguous()), B, H2, W2)
        return ops.nhwc_to_nchw(out, B, 2 * H2, 2 * W2, out.shape[-1])


class _GlobalConvModule(PackedModule):
    def __init__(self, in_dim, out_dim, kernel_size):
        super().__init__()
        pad0 = int((kernel_size[0] - 1) / 2)
        pad1 = int((kernel_size[1] - 1) / 2)
        self.conv_l1 = nn.Conv2d(in_dim, out_dim, kernel_size=(kernel_size[0], 1), padding=(pad0, 0))
        self.conv_l2 = nn.Conv2d(out_dim, out_dim, kernel_size=(1, kernel_size[1]), padding=(0, pad1))
        self.conv_r1 = nn.Conv2d(in_dim, out_dim, kernel_size=(1, kernel_size[1]), padding=(0, pad1))
        self.conv_r2 = nn.Conv2d(out_dim, out_dim, kernel_size=(kernel_size[0], 1), padding=(pad0, 0))

    def nhwc(self, x, B, H, W):
        if ops.tensor_cores() and x.dtype == torch.float32:
            x = ops.cast16(x)                       # one operand copy feeds both branches
        l = _conv(self, "l2", self.conv_l2, _conv(self, "l1", self.conv_l1, x, B, H, W, out16=True), B, H, W)
        r = _conv(self, "r2", self.conv_r2, _conv(self, "r1", self.conv_r1, x, B, H, W, out16=True), B, H, W, residual=l)
        return r

    def forward(self, x):
        require_inference(self)
        B, _, H, W = x.shape
        out = self.nhwc(ops.nchw_to_nhwc(x.contiguous()), B, H, W)
        return ops.nhwc_to_nchw(out, B, H, W, out.shape[-1])


class Decoder(PackedModule):

    def __init__(self, in_channels=2304, out_channels=1, kernel_size=7, num_classes=32, dap_k=2,
                 features=[256, 256, 256, 256, 256], input_token_temporal_dims=[1, 1, 3], rgb_features=[320, 640, 1280, 2560],
                 shape=[56, 28, 14, 7]):
        super().__init__()
        self.input_token_temporal_dims = input_token_temporal_dims
        max_t = max(self.input_token_temporal_dims)
        self.shape = shape
        self.dap_k = dap_k
        nc, k2 = num_classes, num_classes * dap_k ** 2

        def dec(cin):
            return nn.Sequential(nn.Conv2d(cin, k2, 3, padding=1), nn.GroupNorm(num_groups=8, num_channels=k2),
                                 nn.ReLU(inplace=True), nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True))
        self.decoder_2 = dec(nc)
        self.decoder_3 = dec(k2)
        self.decoder_4 = dec(k2)
        self.decoder_5 = dec(k2)
        self.final_out = nn.Conv2d(nc, out_channels, 3, padding=1)

        def rgb(i):
            return nn.Sequential(nn.Conv3d(rgb_features[i], features[i], kernel_size=(max_t, 1, 1), padding=0, stride=(max_t, 1, 1)),
                                 nn.GroupNorm(num_groups=16, num_channels=features[i]), nn.ReLU(inplace=True))
        self.rgb_decoder_1 = rgb(0)
        self.rgb_decoder_2 = rgb(1)
        self.rgb_decoder_3 = rgb(2)
        self.rgb_decoder_4 = rgb(3)
        ks = (kernel_size, kernel_size)
        self.gcm1 = _GlobalConvModule(features[-1] + in_channels, 1 * nc * 4, ks)
        self.gcm2 = _GlobalConvModule(features[-2], 1 * nc, ks)
        self.gcm3 = _GlobalConvModule(features[-3], 1 * k2, ks)
        self.gcm4 = _GlobalConvModule(features[-4], 1 * k2, ks)
        self.ecre = nn.PixelShuffle(2)
        self.seb1 = SEB(features[-1], features[-2])
        self.seb2 = SEB(features[-2] + features[-1], features[-3])
        self.seb3 = SEB(features[-3] + features[-2] + features[-1], features[-4])
        self.upsample2 = nn.Upsample(scale_factor=2, mode="bilinear")
        self.upsample4 = nn.Upsample(scale_factor=4, mode="bilinear")
        self.DAP = nn.Sequential(nn.PixelShuffle(dap_k), nn.AvgPool2d((dap_k, dap_k)))

        def freq(cin, cout, groups):
            return nn.Sequential(nn.AvgPool2d(2, stride=2), nn.Conv2d(cin, cout, 3, padding=1),
                                 nn.GroupNorm(num_groups=groups, num_channels=cout), nn.Sigmoid())
        self.decoder_frequency_0 = freq(9, k2, 8)
        self.decoder_frequency_1 = freq(k2, k2, 8)
        self.decoder_frequency_2 = freq(k2, k2, 8)
        self.decoder_frequency_3 = freq(k2, nc, 4)
        self.decoder_frequency_4 = freq(nc, k2, 8)

    # ---- pieces ----------------------------------------------------------------------------------------------
    def merge_views_along_channel_axis(self, tokens, height):
        """API parity with decoder.py:43-53 (b c t h w); the fused forward never materialises this tensor."""
        max_t = max(self.input_token_temporal_dims)
        xs = []
        for idx, x in enumerate(tokens):
            bs, time, n, c = x.shape
            t = self.input_token_temporal_dims[idx]
            x = x.reshape(bs, t, (time * n) // t, c)
            xs.append(x.repeat((1, max_t // x.shape[1], 1, 1)))
        out = torch.cat(xs, dim=-1)
        b, t, n, c = out.shape
        return out.view(b, t, height, n // height, c).permute(0, 4, 1, 2, 3)

    def _rgb_stage(self, idx, stage_tokens, B, h):
        """merge_views (:43-53) + Conv3d k=s=(3,1,1) (:98-120) + GroupNorm(16) + ReLU as one GEMM over [v1|v2|v3_t0|v3_t1|v3_t2]
        with the t-repeated views' weights pre-summed over t."""
        seq = getattr(self, "rgb_decoder_%d" % (idx + 1))
        conv, gn = seq[0], seq[1]
        T = max(self.input_token_temporal_dims)
        hw = h * h
        widths = [t.shape[-1] for t in stage_tokens]
        Kp = sum(w if self.input_token_temporal_dims[v] == 1 else w * T for v, w in enumerate(widths))

        def make():
            w = conv.weight.detach()[:, :, :, 0, 0]                 # (Cout, Cin, T)
            parts, c0 = [], 0
            for v, wd in enumerate(widths):
                blk = w[:, c0:c0 + wd, :]
                parts.append(blk.sum(-1) if self.input_token_temporal_dims[v] == 1 else blk.permute(0, 2, 1).reshape(w.shape[0], -1))
                c0 += wd
            wp = torch.cat(parts, 1).contiguous()
            return ops.cast16(wp) if ops.tensor_cores() else wp
        wp = self._packed("rgbw%d" % idx, [conv.weight], make)
        a = torch.empty((B * hw, Kp), dtype=ops.act_dtype(), device=conv.weight.device)
        col = 0
        for v, t in enumerate(stage_tokens):
            tok = t.reshape(B, -1, widths[v]).contiguous()
            if self.input_token_temporal_dims[v] == 1:
                ops.gather_rows(tok, widths[v], a, Kp, col, B, hw, hw)
                col += widths[v]
            else:
                for tt in range(T):
                    ops.gather_rows(tok, widths[v], a, Kp, col, B, hw, T * hw, add=tt * hw)
                    col += widths[v]
        y = ops.linear(a, wp, conv.bias)
        return ops.groupnorm_nhwc(y, gn.weight, gn.bias, B, hw, y.shape[-1], gn.num_groups, ops.ACT_RELU, gn.eps).view(B, h, h, -1)

    NCHW_FEATS = False      # see forward(): x_feats as a channels-last view (False) or an explicitly transposed NCHW tensor (True)

    def _freq_stage(self, name, x, B, H, W, pooled=False, ld_in=None):
        """AvgPool2 -> Conv3x3 -> GroupNorm -> Sigmoid (:147-181). x NHWC (B,H,W,C) (already pooled if `pooled`)."""
        seq = getattr(self, name)
        if not pooled:
            x = ops.resample_nhwc(x, B, H, W, x.shape[-1], ops.RS_AVGPOOL2, out_dtype=_conv_in_dtype())
            H, W = H // 2, W // 2
        c = _conv(self, name, seq[1], x, B, H, W, ld_in=ld_in)
        return ops.groupnorm_nhwc(c, seq[2].weight, seq[2].bias, B, H * W, c.shape[-1], seq[2].num_groups, ops.ACT_SIGMOID, seq[2].eps)

    def _dec_stage(self, name, x, B, H, W, dap=False, gate=None, out_dtype=torch.float32):
        """Conv3x3 -> GroupNorm(8) -> ReLU -> Upsample x2 align_corners=True (:67-95); with dap=True the DAP channel-group
        mean (:140-143) is taken before the upsample (they commute)."""
        seq = getattr(self, name)
        c = _conv(self, name, seq[0], x, B, H, W)
        k = self.dap_k ** 2 if dap else 1
        fused = k == 4 and c.shape[-1] % 4 == 0          # the DAP mean of 4 consecutive channels comes out of the GroupNorm apply pass
        g = ops.groupnorm_nhwc(c, seq[1].weight, seq[1].bias, B, H * W, c.shape[-1], seq[1].num_groups, ops.ACT_RELU, seq[1].eps,
                               quad_mean=fused)
        C = g.shape[-1]
        if dap and not fused:
            g = ops.channel_group_mean(g, B * H * W, C, k)
            C //= k
        # gate: the next stage's input is this stage's output times a frequency map (decoder.py:221 `x * freq0`): multiplied inside
        # the upsample kernel instead of a separate pass over the (B, 2H, 2W, C) map
        return ops.resample_nhwc(g, B, H, W, C, ops.RS_UP_ALIGNED, 2, mul=gate, out_dtype=out_dtype)

    def forward(self, x, view_x, ffinfo):
        """x (B,2304,n,n), view_x [4][3] of (B,1,L,C), ffinfo (B,9,S,S) -> (binary_mask (B,1,S,S), x_feats (B,32,S,S))."""
        require_inference(self)
        B = x.shape[0]
        s0, s1, s2, s3 = self.shape
        xh = x.permute(0, 2, 3, 1)
        xh = xh if xh.is_contiguous() else ops.nchw_to_nhwc(x.contiguous())            # (B,n,n,2304)
        S = ffinfo.shape[-1]
        # Independent branches run on lanes (streams.py): lane 3 the frequency pyramid, lanes 0-1 the rgb pyramid levels with
        # their SEB / global-conv modules, the decoder_2..5 chain follows on lane 0.  Hand-overs are events; inputs that another
        # lane produced (stage features, the frequency map) are awaited with need_tensor, final_x with need_main, so that inside
        # an outer region (mumpy_b200.forward) these branches overlap the encoder's tail.  Lane 2 is left to the heaviest view.
        with streams.region(x.device) as reg:
            with reg.lane(3):
                reg.need_tensor(ffinfo)
                # AvgPool2 of decoder_frequency_0 fused into the layout change; pixel stride padded to 16 channels so that the
                # 9-channel map can feed the TMA convolution (the padding is never read)
                cf = ffinfo.shape[1]
                ldf = (cf + 7) // 8 * 8
                # (zero-filled: the operand cast below converts the padding channels too, and stale memory there could trip the
                # fp16 range guard although the convolution never reads them)
                f_in = torch.zeros((B, S // 2, S // 2, ldf), dtype=torch.float32, device=x.device)
                ops.nchw_to_nhwc(ffinfo.contiguous().float(), pool2=True, out=f_in, ld_out=ldf)
                freq0 = self._freq_stage("decoder_frequency_0", f_in, B, S // 2, S // 2, pooled=True, ld_in=ldf)
                reg.publish("freq0", freq0)
                freq1 = self._freq_stage("decoder_frequency_1", freq0, B, S // 2, S // 2)
                reg.publish("freq1", freq1)
                freq2 = self._freq_stage("decoder_frequency_2", freq1, B, S // 4, S // 4)
                reg.publish("freq2", freq2)
                freq3 = self._freq_stage("decoder_frequency_3", freq2, B, S // 8, S // 8)
                reg.publish("freq3", freq3)
                freq4 = self._freq_stage("decoder_frequency_4", freq3, B, S // 16, S // 16)
                reg.publish("freq4", freq4)
            with reg.lane(0):
                reg.need_tensor(*view_x[3])
                rgb4 = self._rgb_stage(3, view_x[3], B, s3)
                reg.publish("rgb4", rgb4)
            with reg.lane(1):
                reg.need_tensor(*view_x[2])
                rgb3 = self._rgb_stage(2, view_x[2], B, s2)
                reg.publish("rgb3", rgb3)
                reg.need_tensor(*view_x[1])
                rgb2 = self._rgb_stage(1, view_x[1], B, s1)
                reg.publish("rgb2", rgb2)
            c2, c3, c4 = rgb2.shape[-1], rgb3.shape[-1], rgb4.shape[-1]

            with reg.lane(1):
                reg.need("rgb4")
                seb1 = self.seb1.nhwc(rgb3, rgb4, B, s3, s3, out_dtype=_conv_in_dtype())
                gcn1 = self.gcm2.nhwc(seb1, B, s2, s2)
                reg.publish("gcn1", gcn1)
                cat2 = torch.empty((B, s2, s2, c3 + c4), dtype=_conv_in_dtype(), device=x.device)  # cat[rgb3, up2(rgb4)] (:210)
                ops.resample_nhwc(rgb3, B, s2, s2, c3, ops.RS_IDENTITY, out=cat2, ld_out=c3 + c4, out_col=0)
                ops.resample_nhwc(rgb4, B, s3, s3, c4, ops.RS_UP_HALFPIX, 2, out=cat2, ld_out=c3 + c4, out_col=c3)
                seb2 = self.seb2.nhwc(rgb2, cat2, B, s2, s2, out_dtype=_conv_in_dtype())
                gcn2 = self.gcm3.nhwc(seb2, B, s1, s1)
                reg.publish("gcn2", gcn2)
            with reg.lane(0):
                reg.need_tensor(*view_x[0])
                rgb1 = self._rgb_stage(0, view_x[0], B, s0)
                reg.need("rgb2")
                reg.need("rgb3")
                ct = c2 + c3 + c4
                cat3 = torch.empty((B, s1, s1, ct), dtype=_conv_in_dtype(), device=x.device)     # cat[rgb2, up2(rgb3), up4(rgb4)] (:213)
                ops.resample_nhwc(rgb2, B, s1, s1, c2, ops.RS_IDENTITY, out=cat3, ld_out=ct, out_col=0)
                ops.resample_nhwc(rgb3, B, s2, s2, c3, ops.RS_UP_HALFPIX, 2, out=cat3, ld_out=ct, out_col=c2)
                ops.resample_nhwc(rgb4, B, s3, s3, c4, ops.RS_UP_HALFPIX, 4, out=cat3, ld_out=ct, out_col=c2 + c3)
                seb3 = self.seb3.nhwc(rgb1, cat3, B, s1, s1, out_dtype=_conv_in_dtype())
                gcn3 = self.gcm4.nhwc(seb3, B, s0, s0)

                # the only part that needs the encoder's final tokens (produced on the caller's stream)
                reg.need_main()
                cin = c4 + xh.shape[-1]
                cat0 = torch.empty((B, s3, s3, cin), dtype=_conv_in_dtype(), device=x.device)    # cat[rgb4, x] (:204)
                ops.resample_nhwc(rgb4, B, s3, s3, c4, ops.RS_IDENTITY, out=cat0, ld_out=cin, out_col=0)
                ops.resample_nhwc(xh, B, s3, s3, xh.shape[-1], ops.RS_IDENTITY, out=cat0, ld_out=cin, out_col=c4)
                gcn0 = self.gcm1.nhwc(cat0, B, s3, s3)
                reg.need("freq4")
                g0 = ops.mul_add(gcn0, freq4)
                out1 = ops.resample_nhwc(g0, B, s3, s3, g0.shape[-1], ops.RS_PIXEL_SHUFFLE2)  # ecre (:205)

                reg.need("gcn1")
                reg.need("freq3")
                d = self._dec_stage("decoder_2", ops.mul_add(gcn1, freq3, out1, out_dtype=_conv_in_dtype()), B, s2, s2)        # (:218)
                reg.need("gcn2")
                reg.need("freq2")
                d = self._dec_stage("decoder_3", ops.mul_add(gcn2, freq2, d, out_dtype=_conv_in_dtype()), B, s1, s1)           # (:219)
                reg.need("freq1")
                reg.need("freq0")
                d = self._dec_stage("decoder_4", ops.mul_add(gcn3, freq1, d, out_dtype=_conv_in_dtype()), B, s0, s0, gate=freq0, out_dtype=_conv_in_dtype())   # (:220) and the `* freq0` of (:221)
                feats = self._dec_stage("decoder_5", d, B, 2 * s0, 2 * s0, dap=True)                    # (:221-222) (B,S,S,32)
            reg.wait_lanes()          # hand the result back to the caller's stream (the lanes stay forked in a nested region)
        Sf = 4 * s0
        mask = _conv(self, "final_out", self.final_out, feats, B, Sf, Sf)                  # (B,S,S,1) == (B,1,S,S)
        # x_feats (B,32,S,S): the reference's second return value, which test.py discards (test.py:95).  The kernels produce it
        # pixel-major; it is handed back as a channels-last VIEW with the reference's shape and values (strides differ:
        # `.contiguous()` gives the reference's memory layout).  Decoder.NCHW_FEATS = True restores the explicit transposition
        # kernel (205 MB in, 205 MB out per 32 clips: 130 us of a 14 ms step for a tensor nobody reads).
        if self.NCHW_FEATS:
            x_feats = ops.nhwc_to_nchw(feats, B, Sf, Sf, feats.shape[-1])
        else:
            x_feats = feats.view(B, Sf, Sf, feats.shape[-1]).permute(0, 3, 1, 2)
        return mask.view(B, self.final_out.out_channels, Sf, Sf) if self.final_out.out_channels == 1 else \
            ops.nhwc_to_nchw(mask, B, Sf, Sf, mask.shape[-1]), x_feats
