"""GPU: the 16-bit-only / 16-bit-output kernels through the C ABI against the oracle's arithmetic on the SAME rounded operands
(round-1 review item: these kernels were only covered by the end-to-end tolerances).

  cva_offsets_reg_kernel / cva_offsets_kernel (fallback for wide groups), cva_sample_kernel<16,16>, cva_attention_mma_kernel,
  mha3_kernel<16>, every layernorm_vec_kernel instantiation with 16-bit output, MergeRows LayerNorm, resample_vec_kernel modes,
  gather_rows_vec_kernel.  Window subsets of the configs[1] (batch-64) shapes are compared where the oracle would be slow.
"""
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu
DTS = [torch.bfloat16, torch.float16]


def _ops():
    import mumpy_b200
    return mumpy_b200.ops


def _ulp(dt):
    return 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11


# ------------------------------------------------------------------------------------------------ deformable attention
def _offset_pixels(q, dw_w, dw_b, ln_g, ln_b, pw, N1, ws, groups):
    """deformableAttention.py:253-258,334-340,353-356 on q (N1, P, C): pixel coordinates (N1, groups, P, 2) = (y, x)."""
    C = q.shape[-1]
    Cg = C // groups
    qmap = q.reshape(N1, ws, ws, groups, Cg)
    padded = torch.zeros(N1, ws + 4, ws + 4, groups, Cg)
    padded[:, 2:2 + ws, 2:2 + ws] = qmap
    acc = torch.zeros_like(qmap)
    for a in range(5):
        for b in range(5):
            acc = acc + padded[:, a:a + ws, b:b + ws] * dw_w.view(Cg, 5, 5)[:, a, b]
    acc = orc.gelu(orc.layer_norm(acc + dw_b, ln_g, ln_b))
    off = torch.tanh(acc @ pw.view(2, Cg).t()) * (1.0 / ws) * 2.0
    ref = (torch.arange(ws, dtype=torch.float32) + 0.5) / ws * 2.0 - 1.0
    py = ((off[..., 0] + ref.view(1, ws, 1, 1) + 1.0) * 0.5 * (ws - 1)).permute(0, 3, 1, 2).reshape(N1, groups, ws * ws)
    px = ((off[..., 1] + ref.view(1, 1, ws, 1) + 1.0) * 0.5 * (ws - 1)).permute(0, 3, 1, 2).reshape(N1, groups, ws * ws)
    return torch.stack([py, px], -1)


@pytest.mark.parametrize("C,ws,B,TH,W", [(96, 7, 2, 14, 14), (384, 7, 1, 14, 14), (768, 7, 3, 7, 7), (192, 8, 1, 16, 16), (768, 8, 1, 8, 8)])
def test_cva_offsets_kernels(C, ws, B, TH, W):
    """Group widths 32 / 128 / 64 run the register-resident kernel, 256 (C = 768, stage 3 of views 1 and 2) the CTA-per-unit
    fallback: both against the oracle's offset network."""
    ops = _ops()
    groups, Cg = 3, C // 3
    q = util.seeded_input((B, TH * W, C), 1)
    dw_w, dw_b = util.seeded_input((Cg, 25), 2) / 5.0, util.seeded_input((Cg,), 3) * 0.1
    ln_g, ln_b = 1.0 + 0.1 * util.seeded_input((Cg,), 4), 0.1 * util.seeded_input((Cg,), 5)
    pw = util.seeded_input((2, Cg), 6) / Cg ** 0.5
    N1 = B * (TH // ws) * (W // ws)
    qw = orc.window_partition(q.view(B, TH, W, C), ws).reshape(N1, ws * ws, C)
    ref = _offset_pixels(qw, dw_w, dw_b, ln_g, ln_b, pw, N1, ws, groups)
    pix = ops.cva_offsets(q.cuda(), dw_w.cuda(), dw_b.cuda(), ln_g.cuda(), ln_b.cuda(), pw.cuda(), B, TH, W, C, groups, ws)
    assert pix.shape == ref.shape
    assert util.maxabs(pix, ref) < 2e-5


def _pairing(N1, N2, nW1, per_clip):
    r = N2 // N1
    j = torch.arange(N2)
    if not per_clip:
        return j % N1
    i_out = j // r
    clip = i_out // nW1
    return clip * nW1 + ((i_out % nW1) * r + j % r) % nW1


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("per_clip", [False, True])
@pytest.mark.parametrize("B,H,T2,C", [(2, 14, 3, 96), (3, 7, 1, 384), (64, 56, 3, 96)])
def test_cva_sample_16bit(B, H, T2, C, per_clip, dt):
    """cva_sample_kernel<16-bit in, 16-bit out>: bilinear sampling of the key/value canvas at the paired query window's
    offsets (deformableAttention.py:329-330,353-356) vs the oracle on the same rounded canvas; the result is rounded once.
    (64, 56, 3, 96) is the configs[1] shape (4096 x 3 query windows, 12288 kv windows): a window subset is compared."""
    ops = _ops()
    ws, groups, W = 7, 3, H
    TH1, TH2 = H, T2 * H
    nW1 = (TH1 // ws) * (W // ws)
    N1, N2 = B * nW1, B * (TH2 // ws) * (W // ws)
    g = torch.Generator().manual_seed(11)
    x2 = torch.randn((B, TH2 * W, C), generator=g).to(dt)
    pix = torch.rand((N1, groups, ws * ws, 2), generator=g) * 8.0 - 1.0          # incl. positions outside [0, 6]: zero padding
    out = ops.cva_sample(x2.cuda(), pix.cuda(), B, TH1, TH2, W, C, groups, ws, per_clip, dt)
    assert out.dtype == dt and out.shape == (N2 * ws * ws, C)
    qidx = _pairing(N1, N2, nW1, per_clip)
    sel = torch.arange(N2) if N2 <= 64 else torch.cat([torch.arange(0, 40), torch.arange(N2 // 2 - 20, N2 // 2 + 20), torch.arange(N2 - 40, N2)])
    x2w = orc.window_partition(x2.float().view(B, TH2, W, C), ws).reshape(N2, ws, ws, groups, C // groups)[sel]
    img = x2w.permute(0, 3, 4, 1, 2).reshape(len(sel) * groups, C // groups, ws, ws)
    p = pix[qidx[sel]].reshape(len(sel) * groups, ws * ws, 2)
    ref = orc.bilinear_sample_zeros(img, p[..., 0], p[..., 1]).reshape(len(sel), C, ws * ws).transpose(1, 2)
    got = out.view(N2, ws * ws, C)[sel.cuda()].float().cpu()
    assert util.maxabs(got, ref) <= _ulp(dt) * max(1.0, float(ref.abs().max())) + 1e-5


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("per_clip", [False, True])
@pytest.mark.parametrize("B,H,T2,C,heads,ws", [(2, 14, 3, 96, 3, 7), (2, 7, 1, 384, 12, 7), (1, 16, 3, 64, 2, 8)])
def test_cva_attention_16bit(B, H, T2, C, heads, ws, per_clip, dt):
    """cva_attention_mma_kernel: softmax(q k^T / sqrt(d)) v per (kv window, head), queries of the paired window, summed over
    the temporal ratio (deformableAttention.py:360-364,390-395) vs the oracle on the rounded q / kv (P rounded to 16 bits
    before PV inside the kernel: 2e-2 on O(1) outputs)."""
    ops = _ops()
    W, TH1, TH2, P = H, H, T2 * H, ws * ws
    nW1 = (TH1 // ws) * (W // ws)
    N1, N2 = B * nW1, B * (TH2 // ws) * (W // ws)
    r, d = N2 // N1, C // heads
    q = util.seeded_input((B, TH1 * W, C), 21)
    kv = util.seeded_input((N2 * P, 2 * C), 22).to(dt)
    out = ops.cva_attention(q.cuda(), kv.cuda(), B, TH1, TH2, W, C, heads, ws, per_clip)
    qw = orc.window_partition(q.to(dt).float().view(B, TH1, W, C), ws).reshape(N1, P, C)
    qidx = _pairing(N1, N2, nW1, per_clip)
    k, v = kv.float().view(N2, P, 2 * C)[..., :C], kv.float().view(N2, P, 2 * C)[..., C:]
    qh = qw[qidx].reshape(N2, P, heads, d).permute(0, 2, 1, 3)
    kh, vh = k.reshape(N2, P, heads, d).permute(0, 2, 1, 3), v.reshape(N2, P, heads, d).permute(0, 2, 1, 3)
    o = (torch.softmax(qh @ kh.transpose(-2, -1) * d ** -0.5, -1) @ vh).permute(0, 2, 1, 3).reshape(N1, r, P, C).sum(1)
    assert out.dtype == dt and out.shape == (N1 * P, C)
    assert util.maxabs(out.float().view(N1, P, C), o) < 2e-2 * r


# ------------------------------------------------------------------------------------------------ global ViT attention (N = 3)
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("Bn,C,heads", [(98, 768, 12), (1568, 768, 12), (5, 128, 2)])
def test_mha_short_16bit(Bn, C, heads, dt):
    """mha3_kernel: attention over the 3 temporal tokens per (clip, position) (blocks.py:56-71 under vmap,
    multiTemporalViewEncoder.py:741), scale applied after q k^T."""
    ops = _ops()
    N, d = 3, C // heads
    qkv = util.seeded_input((Bn, N, 3 * C), 31).to(dt)
    out = ops.mha_short(qkv.cuda(), Bn, N, C, heads)
    q, k, v = qkv.float().view(Bn, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax((q @ k.transpose(-2, -1)) * d ** -0.5, -1) @ v).transpose(1, 2).reshape(Bn, N, C)
    assert out.dtype == dt
    assert util.maxabs(out.float(), ref) <= 2 * _ulp(dt) * max(1.0, float(ref.abs().max()))


# ------------------------------------------------------------------------------------------------ LayerNorm family
@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("C", [96, 128, 192, 256, 384, 512, 768, 1024, 36])
def test_layernorm_16bit_every_width(C, dt):
    """One case per layernorm_vec_kernel instantiation (C/4 = 24 ... 256 vectors; C = 36 takes the generic kernel), ragged row
    count, rows with a large common offset (two-pass statistics)."""
    ops = _ops()
    rows = 1000 + 7
    x = util.seeded_input((rows, C), 41) * 3.0 + util.seeded_input((rows, 1), 42) * 20.0
    gam, bet = 1.0 + 0.2 * util.seeded_input((C,), 43), 0.2 * util.seeded_input((C,), 44)
    ref = orc.layer_norm(x.double(), gam.double(), bet.double()).float()
    out = ops.layernorm(x.cuda(), gam.cuda(), bet.cuda(), 1e-5, out_dtype=dt)
    assert out.dtype == dt
    assert util.maxabs(out.float(), ref) <= _ulp(dt) * float(ref.abs().max()) + 1e-4
    out32 = ops.layernorm(x.cuda(), gam.cuda(), bet.cuda(), 1e-5, out_dtype=torch.float32)
    assert util.maxabs(out32, ref) < 2e-5


@pytest.mark.parametrize("dt", DTS)
@pytest.mark.parametrize("B,TH,W,C", [(2, 28, 28, 96), (1, 42, 14, 128), (3, 14, 14, 384), (1, 8, 8, 64)])
def test_patch_merge_norm_16bit(B, TH, W, C, dt):
    """MergeRows LayerNorm: 2x2 neighbourhood gather [x(0,0), x(1,0), x(0,1), x(1,1)] + LN(4C) (swinTransformer.py:357-364)."""
    ops = _ops()
    x = util.seeded_input((B, TH * W, C), 51)
    gam, bet = 1.0 + 0.2 * util.seeded_input((4 * C,), 52), 0.2 * util.seeded_input((4 * C,), 53)
    xx = x.view(B, TH, W, C)
    cat = torch.cat([xx[:, 0::2, 0::2], xx[:, 1::2, 0::2], xx[:, 0::2, 1::2], xx[:, 1::2, 1::2]], -1).reshape(B, -1, 4 * C)
    ref = orc.layer_norm(cat, gam, bet)
    out = ops.patch_merge_norm(x.cuda(), gam.cuda(), bet.cuda(), B, TH, W, C, 1e-5, out_dtype=dt)
    assert out.dtype == dt and out.shape == ref.shape
    assert util.maxabs(out.float(), ref) <= _ulp(dt) * float(ref.abs().max()) + 1e-4


# ------------------------------------------------------------------------------------------------ decoder data movement
@pytest.mark.parametrize("mode,scale", [("up_aligned", 2), ("up_halfpix", 2), ("up_halfpix", 4), ("avgpool", 2), ("shuffle", 2), ("identity", 1)])
def test_resample_modes(mode, scale):
    """resample_vec_kernel: nn.Upsample(bilinear) with both align_corners settings (decoder.py:10,72,136-137), AvgPool2d(2),
    PixelShuffle(2), with the fused gate (* mul) and skip (+ add) and a strided write into a wider concat buffer."""
    ops = _ops()
    B, H, W, C = 2, 14, 10, 64
    x = util.seeded_input((B, C, H, W), 61)
    code = {"up_aligned": ops.RS_UP_ALIGNED, "up_halfpix": ops.RS_UP_HALFPIX, "avgpool": ops.RS_AVGPOOL2, "shuffle": ops.RS_PIXEL_SHUFFLE2,
            "identity": ops.RS_IDENTITY}[mode]
    if mode.startswith("up"):
        ref = orc.upsample_bilinear(x, scale, mode == "up_aligned")
    elif mode == "avgpool":
        ref = orc.avg_pool2(x)
    elif mode == "shuffle":
        ref = orc.pixel_shuffle2(x)
    else:
        ref = x
    ref = ref.permute(0, 2, 3, 1).contiguous()
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.resample_nhwc(xn, B, H, W, C, code, scale)
    assert out.shape == ref.shape and util.maxabs(out, ref) < 1e-5
    mul, add = util.seeded_input(tuple(ref.shape), 62), util.seeded_input(tuple(ref.shape), 63)
    out2 = ops.resample_nhwc(xn, B, H, W, C, code, scale, mul=mul.cuda(), add=add.cuda())
    assert util.maxabs(out2, ref * mul + add) < 1e-5
    Co = ref.shape[-1]
    wide = torch.full(tuple(ref.shape[:-1]) + (Co + 32,), -7.0).cuda()
    ops.resample_nhwc(xn, B, H, W, C, code, scale, out=wide, ld_out=Co + 32, out_col=16)
    assert util.maxabs(wide[..., 16:16 + Co], ref) < 1e-5
    assert bool((wide[..., :16] == -7.0).all()) and bool((wide[..., 16 + Co:] == -7.0).all())


@pytest.mark.parametrize("dt", DTS + [torch.float32])
def test_gather_rows_view_merge(dt):
    """gather_rows_vec_kernel as the encoder uses it (merge_views_along_channel_axis, multiTemporalViewEncoder.py:710-718): rows
    ordered (b, n, t); views with one temporal token are repeated over t, view 3's token (t, n) is row t*49 + n."""
    ops = _ops()
    B, n, T = 3, 49, 3
    v1, v3 = util.seeded_input((B, n, 96), 71), util.seeded_input((B, n * T, 128), 72)
    merged = torch.zeros((B * n * T, 96 + 128), dtype=dt).cuda()
    ops.gather_rows(v1.cuda(), 96, merged, 224, 0, B, n * T, n, div=T, mul_hi=1, mul_lo=0)
    ops.gather_rows(v3.cuda(), 128, merged, 224, 96, B, n * T, n * T, div=T, mul_hi=1, mul_lo=n)
    ref = torch.cat([v1.unsqueeze(2).expand(B, n, T, 96), v3.view(B, T, n, 128).permute(0, 2, 1, 3)], -1).reshape(B * n * T, 224)
    assert torch.equal(merged.float().cpu(), ref.to(dt).float())


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_resample_and_mul_add_write_operands_directly(dt):
    """16-bit outputs of resample_vec_kernel (all vector modes, strided into a concat buffer, with gate and skip) and of
    mul_add_vec_kernel equal the fp32 result passed through mumpy_cast16, bit for bit: the decoder writes convolution operands
    straight from the producing kernel."""
    import mumpy_b200
    ops = mumpy_b200.ops
    B, H, W, C = 2, 14, 10, 32
    x = util.seeded_input((B, H, W, C), 1).cuda()
    for mode, scale, Ho, Wo in ((ops.RS_IDENTITY, 1, H, W), (ops.RS_UP_ALIGNED, 2, 2 * H, 2 * W), (ops.RS_UP_HALFPIX, 2, 2 * H, 2 * W),
                                (ops.RS_UP_HALFPIX, 4, 4 * H, 4 * W), (ops.RS_AVGPOOL2, 2, H // 2, W // 2)):
        mul = util.seeded_input((B, Ho, Wo, C), 2).cuda()
        add = util.seeded_input((B, Ho, Wo, C), 3).cuda()
        ref = ops.cast16(ops.resample_nhwc(x, B, H, W, C, mode, scale, mul=mul, add=add), dt)
        out = ops.resample_nhwc(x, B, H, W, C, mode, scale, mul=mul, add=add, out_dtype=dt)
        assert out.dtype == dt and torch.equal(out, ref)
        cat = torch.zeros((B, Ho, Wo, C + 8), dtype=dt, device=x.device)
        ops.resample_nhwc(x, B, H, W, C, mode, scale, out=cat, ld_out=C + 8, out_col=8)
        assert torch.equal(cat[..., 8:], ops.cast16(ops.resample_nhwc(x, B, H, W, C, mode, scale), dt)) and float(cat[..., :8].float().abs().max()) == 0.0
    a, b, c = (util.seeded_input((B, H, W, C), s).cuda() for s in (4, 5, 6))
    assert torch.equal(ops.mul_add(a, b, c, out_dtype=dt), ops.cast16(ops.mul_add(a, b, c), dt))
    assert torch.equal(ops.mul_add(a, b, out_dtype=dt), ops.cast16(ops.mul_add(a, b), dt))
