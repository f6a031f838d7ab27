"""GPU: individual bf16-mode kernels through the C ABI against the oracle on bf16-rounded operands."""
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def _ops():
    import mumpy_b200
    return mumpy_b200.ops


def _pack_conv(w, Cin, dt=torch.bfloat16):
    Cout, _, kh, kw = w.shape
    cb = (Cin + 63) // 64
    wp = torch.zeros(Cout, kh * kw, cb * 64)
    wp[:, :, :Cin] = w.permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)
    return wp.reshape(Cout, -1).to(dt).contiguous()


@pytest.fixture
def pair_mode(request):
    ops = _ops()
    ops.set_gemm_pair_mode(request.param)
    yield request.param
    ops.set_gemm_pair_mode(0)


@pytest.mark.parametrize("pair_mode", [0, 2], indirect=True)
@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw", [
    (1, 7, 7, 256, 128, 3, 3),        # < 128 KB tensor (driver work-around path)
    (2, 14, 14, 32, 128, 3, 3),       # Cin < 64: channel block zero-filled by TMA
    (2, 28, 28, 768, 256, 3, 3),      # seb3
    (2, 56, 56, 256, 128, 7, 1),      # gcm4 conv_l1
    (2, 56, 56, 128, 128, 1, 7),      # gcm4 conv_l2
    (1, 7, 7, 2560, 128, 7, 1),       # gcm1
    (3, 112, 112, 128, 128, 3, 3),    # decoder_5, M tail across image borders
    (5, 7, 7, 128, 64, 3, 3),         # 245 pixels: tiles straddle images
])
def test_conv_implicit_gemm_tcgen05(B, H, W, Cin, Cout, kh, kw, pair_mode):
    ops = _ops()
    x = util.seeded_input((B, Cin, H, W), 1).bfloat16()
    w = (util.seeded_input((Cout, Cin, kh, kw), 2) / (Cin * kh * kw) ** 0.5).bfloat16()
    bias = util.seeded_input((Cout,), 3)
    ref = orc.conv2d(x.float(), w.float(), bias, ((kh - 1) // 2, (kw - 1) // 2)).permute(0, 2, 3, 1)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = ops.conv2d_nhwc_bf16(xn, _pack_conv(w.float(), Cin).cuda(), bias.cuda(), B, H, W, Cin, Cout, kh, kw, (kh - 1) // 2, (kw - 1) // 2)
    assert util.maxabs(out, ref) < 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,W,Cin,Cout,kh,kw", [(4, 7, 7, 2560, 128, 1, 7), (32, 7, 7, 640, 128, 7, 1), (2, 14, 14, 512, 256, 3, 3)])
def test_conv_split_k_matches_single_pass(B, H, W, Cin, Cout, kh, kw, dt):
    """Small maps with a long reduction run split-K (partials in a caller-owned workspace + a reduce kernel that applies bias /
    residual / output rounding): same result as the one-pass kernel up to fp32 summation order, and vs the oracle."""
    ops = _ops()
    x = util.seeded_input((B, Cin, H, W), 1).to(dt)
    w = (util.seeded_input((Cout, Cin, kh, kw), 2) / (Cin * kh * kw) ** 0.5).to(dt)
    bias, res = util.seeded_input((Cout,), 3), util.seeded_input((B, H, W, Cout), 4)
    ref = orc.conv2d(x.float(), w.float(), bias, ((kh - 1) // 2, (kw - 1) // 2)).permute(0, 2, 3, 1) + res
    xn, wp = x.permute(0, 2, 3, 1).contiguous().cuda(), _pack_conv(w.float(), Cin, dt).cuda()
    args = (xn, wp, bias.cuda(), B, H, W, Cin, Cout, kh, kw, (kh - 1) // 2, (kw - 1) // 2)
    one = ops.conv2d_nhwc_bf16(*args, residual=res.cuda(), split_k=False)
    two = ops.conv2d_nhwc_bf16(*args, residual=res.cuda(), split_k=True)
    scale = max(1.0, float(ref.abs().max()))
    assert util.maxabs(two, ref) < 2e-4 * scale
    assert util.maxabs(two, one) < 1e-4 * scale          # fp32 summation order over K = 17920
    lo1 = ops.conv2d_nhwc_bf16(*args, out_dtype=dt, split_k=False)
    lo2 = ops.conv2d_nhwc_bf16(*args, out_dtype=dt, split_k=True)
    assert lo2.dtype == dt and util.maxabs(lo2.float(), lo1.float()) <= (2 ** -7 if dt == torch.bfloat16 else 2 ** -10) * scale


def test_conv_reads_channel_slice_of_wider_map():
    ops = _ops()
    B, H, W, Cin, ld, Cout = 2, 14, 14, 64, 192, 32
    x = util.seeded_input((B, H, W, ld), 1).bfloat16()
    w = (util.seeded_input((Cout, Cin, 3, 3), 2) / (Cin * 9) ** 0.5).bfloat16()
    ref = orc.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), w.float(), None, (1, 1)).permute(0, 2, 3, 1)
    out = ops.conv2d_nhwc_bf16(x.cuda(), _pack_conv(w.float(), Cin).cuda(), None, B, H, W, Cin, Cout, 3, 3, 1, 1, ld_in=ld)
    assert util.maxabs(out, ref) < 2e-4 * max(1.0, float(ref.abs().max()))


def test_conv_odd_channel_count_with_padded_pixel_stride():
    """decoder_frequency_0: 9 input channels at a pixel stride of 16 -- the tensor map's channel extent is 9, TMA zero-fills
    the rest of the 64-channel block, the (garbage) padding channels are never read."""
    ops = _ops()
    B, H, W, Cin, ld, Cout = 2, 28, 28, 9, 16, 128
    x = util.seeded_input((B, H, W, ld), 1).bfloat16()
    x[..., Cin:] = float("nan")
    w = (util.seeded_input((Cout, Cin, 3, 3), 2) / (Cin * 9) ** 0.5).bfloat16()
    ref = orc.conv2d(x[..., :Cin].float().permute(0, 3, 1, 2), w.float(), None, (1, 1)).permute(0, 2, 3, 1)
    out = ops.conv2d_nhwc_bf16(x.cuda(), _pack_conv(w.float(), Cin).cuda(), None, B, H, W, Cin, Cout, 3, 3, 1, 1, ld_in=ld)
    assert util.maxabs(out, ref) < 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,Cin,H,W", [(2, 32, 24, 20), (3, 32, 22, 19), (1, 16, 5, 7), (2, 64, 9, 12), (1, 32, 224, 224)])
def test_conv_cout1(B, Cin, H, W):
    """final_out (decoder.py:95): 3x3, one output channel.  Row-tiled kernel for 16 / 32 input channels (incl. heights that are
    not a multiple of the four-row tile and one-pixel borders), per-pixel kernel otherwise."""
    ops = _ops()
    x = util.seeded_input((B, Cin, H, W), 1)
    w = util.seeded_input((1, Cin, 3, 3), 2) / 17.0
    b = util.seeded_input((1,), 3)
    ref = orc.conv2d(x, w, b, (1, 1))
    out = ops.conv2d_nhwc_cout1(x.permute(0, 2, 3, 1).contiguous().cuda(), w.permute(0, 2, 3, 1).contiguous().cuda(), b.cuda(), B, H, W, Cin, 3, 3, 1, 1)
    assert util.maxabs(out.view(B, 1, H, W), ref) < 1e-5


@pytest.fixture
def attention_tc(request):
    """Table-mode kernel for one test: True = tcgen05 / TMEM (default), False = per-warp mma.sync."""
    ops = _ops()
    ops.set_attention_tc(request.param)
    yield request.param
    ops.set_attention_tc(True)


@pytest.mark.parametrize("attention_tc", [True, False], indirect=True)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,T,H,C,heads,shift,ws", [(2, 3, 14, 64, 2, 3, 7), (2, 1, 14, 96, 3, 0, 7), (1, 3, 7, 128, 4, 0, 7), (3, 3, 28, 128, 4, 3, 7),
                                                    (1, 1, 7, 32, 1, 0, 7), (2, 3, 16, 64, 2, 4, 8), (1, 1, 8, 96, 3, 0, 8)])
def test_window_attention_tensor_pipe(B, T, H, C, heads, shift, ws, dt, attention_tc):
    """16-bit qkv -> tensor-pipe kernels (tcgen05 two-windows-per-accumulator kernel in table mode, mma.sync otherwise) vs the
    oracle's attention on the same rounded qkv (P is rounded to 16 bits before PV inside the kernels: tolerance 2e-2 on O(1)
    outputs).  Covers odd window counts (an unpaired last window), one-window inputs, odd head counts and window size 8."""
    ops = _ops()
    TH, W, N = T * H, H, ws * ws
    qkv = util.seeded_input((B, TH * W, 3 * C), 1).to(dt)
    table = 0.5 * util.seeded_input(((2 * ws - 1) ** 2, heads), 2)
    bias = orc.relative_position_bias(table, ws)
    mask = orc.shifted_window_mask(TH, W, ws, shift) if shift else None
    x = qkv.float().view(B, TH, W, 3 * C)
    if shift:
        x = torch.roll(x, (-shift, -shift), (1, 2))
    xw = orc.window_partition(x, ws)
    Bn = xw.shape[0]
    q, k, v = xw.reshape(Bn, N, 3, heads, 32).permute(2, 0, 3, 1, 4)
    attn = (q * 32 ** -0.5) @ k.transpose(-2, -1) + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(Bn // nW, nW, heads, N, N) + mask.view(1, nW, 1, N, N)).view(Bn, heads, N, N)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(Bn, N, C)
    o = orc.window_reverse(o, ws, TH, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    out = ops.window_attention(qkv.cuda(), bias.cuda(), None if mask is None else mask.cuda(), B, TH, W, C, heads, ws, shift)
    assert out.dtype == dt
    assert util.maxabs(out.float(), o.reshape(B, TH * W, C)) < 2e-2
    # fast path: bias looked up in the raw table, standard shift mask recomputed from region ids
    out2 = ops.window_attention(qkv.cuda(), bias.cuda(), None if mask is None else mask.cuda(), B, TH, W, C, heads, ws, shift,
                                rel_table=table.cuda().contiguous(), standard_mask=mask is not None)
    assert util.maxabs(out2.float(), o.reshape(B, TH * W, C)) < 2e-2
    assert util.maxabs(out2.float(), out.float()) < 1e-2


def test_fast_gelu_epilogue_accuracy():
    """bf16-mode GEMM epilogue GELU (log2-erfc polynomial) vs exact erf GELU, fp32 output."""
    ops = _ops()
    M, N, K = 256, 128, 64
    a = torch.zeros(M, K)
    a[:, 0] = torch.linspace(-8, 8, M)
    w = torch.zeros(N, K)
    w[:, 0] = torch.linspace(0.5, 1.0, N)
    ref = orc.gelu(a.bfloat16().float() @ w.bfloat16().float().t())
    out = ops.linear(a.bfloat16().cuda(), w.bfloat16().cuda(), act=ops.ACT_GELU)
    assert util.maxabs(out, ref) < 2e-6


@pytest.mark.parametrize("mode", ["bf16", "fp16"])
@pytest.mark.parametrize("size,B", [(224, 2), (56, 3)])
def test_faf_tensor_core_path(mode, size, B):
    """FAF in the 16-bit modes: four tcgen05 GEMMs on hi/lo-split operands vs the fp32 oracle (dct.py:71-79).
    Split arithmetic keeps ~16 (bf16) / ~22 (f16) mantissa bits: tolerance 1e-4 / 2e-5 on outputs of magnitude ~3."""
    import mumpy_b200
    from mumpy_b200.models.modules.dct import FAF
    x = util.seeded_input((B, 3, 3, size, size), 17)
    ref = orc.faf_middle(x)
    mumpy_b200.set_precision(mode)
    try:
        y = FAF(size).eval().frame(x.cuda(), 1)
    finally:
        mumpy_b200.set_precision(mumpy_b200.ops.DEFAULT_PRECISION)
    assert y.shape == ref.shape
    assert util.maxabs(y, ref) < (1e-4 if mode == "bf16" else 2e-5)


def test_groupnorm_quad_mean_is_the_dap_of_the_plain_output():
    """decoder_5 tail: GroupNorm -> ReLU -> DAP (mean over each 4 consecutive channels, decoder.py:140-143) in one apply pass
    equals the two-kernel form (GroupNorm map, then channel_group_mean)."""
    ops = _ops()
    B, H, W, C, groups = 2, 12, 10, 128, 8
    x = util.seeded_input((B, H, W, C), 1).cuda()
    gam, bet = util.seeded_input((C,), 2).cuda(), util.seeded_input((C,), 3).cuda()
    full = ops.groupnorm_nhwc(x, gam, bet, B, H * W, C, groups, ops.ACT_RELU)
    ref = ops.channel_group_mean(full, B * H * W, C, 4).view(B, H, W, C // 4)
    ref2 = torch.relu(torch.nn.functional.group_norm(x.cpu().permute(0, 3, 1, 2), groups, gam.cpu(), bet.cpu())).permute(0, 2, 3, 1)
    ref2 = ref2.reshape(B, H, W, C // 4, 4).mean(-1)
    out = ops.groupnorm_nhwc(x, gam, bet, B, H * W, C, groups, ops.ACT_RELU, quad_mean=True)
    assert out.shape == (B, H, W, C // 4)
    assert util.maxabs(out, ref) < 1e-6 and util.maxabs(out, ref2) < 1e-5
