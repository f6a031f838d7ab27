"""GPU: the GEMM kernels through the C ABI against the oracle's linear() on the same seeded operands."""
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K) drawn from the model: qkv / proj / fc1 / fc2 / merge / embed, with ragged M and K tails
    (49, 96, 96), (3136, 288, 96), (3136, 384, 96), (3136, 96, 384), (9408, 384, 128), (2352, 1024, 256),
    (588, 1536, 512), (588, 512, 2048), (147, 3072, 1024), (147, 768, 2560), (1000, 256, 576), (130, 64, 72), (1344, 224, 672),
]


def _ops():
    import mumpy_b200
    return mumpy_b200.ops


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", ["plain", "bias_gelu", "bias_residual"])
def test_linear_fp32_exact(M, N, K, epi):
    ops = _ops()
    a, w = util.seeded_input((M, K), 1), util.seeded_input((N, K), 2) / K ** 0.5
    bias = util.seeded_input((N,), 3) if epi != "plain" else None
    res = util.seeded_input((M, N), 4) if epi == "bias_residual" else None
    ref = orc.linear(a, w, bias)
    if epi == "bias_gelu":
        ref = orc.gelu(ref)
    if res is not None:
        ref = ref + res
    out = ops.linear(a.cuda(), w.cuda(), None if bias is None else bias.cuda(), None if res is None else res.cuda(),
                     act=ops.ACT_GELU if epi == "bias_gelu" else ops.ACT_NONE)
    assert util.maxabs(out, ref) < 2e-5


@pytest.fixture
def pair_mode(request):
    """CTA-pair policy of the tensor-core GEMM for one test (0 = 1-CTA tiles, 2 = cta_group::2 pairs whenever legal)."""
    ops = _ops()
    ops.set_gemm_pair_mode(request.param)
    yield request.param
    ops.set_gemm_pair_mode(4)          # the library default


@pytest.mark.parametrize("pair_mode", [0, 2], indirect=True)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", ["plain", "bias_gelu_16out", "bias_residual"])
def test_linear_tcgen05_16bit(M, N, K, epi, dt, pair_mode):
    """bf16 / f16 operands, fp32 accumulation: compare with the fp32 oracle on the rounded operands.
    Tolerance: products are exact in fp32, so only summation order differs (<= ~K * 2^-24 relative) plus one rounding of
    the result when the output is 16-bit (2^-9 relative for bf16, 2^-12 for f16)."""
    ops = _ops()
    a = util.seeded_input((M, K), 1).to(dt)
    w = (util.seeded_input((N, K), 2) / K ** 0.5).to(dt)
    bias = util.seeded_input((N,), 3) if epi != "plain" else None
    res = util.seeded_input((M, N), 4) if epi == "bias_residual" else None
    ref = orc.linear(a.float(), w.float(), bias)
    if epi == "bias_gelu_16out":
        ref = orc.gelu(ref)
    if res is not None:
        ref = ref + res
    out16 = epi == "bias_gelu_16out"
    out = ops.linear(a.cuda(), w.cuda(), None if bias is None else bias.cuda(), None if res is None else res.cuda(),
                     act=ops.ACT_GELU if out16 else ops.ACT_NONE, out_dtype=dt if out16 else torch.float32)
    tol = (2e-2 if dt == torch.bfloat16 else 3e-3) if out16 else 1e-4
    assert util.maxabs(out.float(), ref) < tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(3136, 96, 96), (588, 512, 512), (130, 128, 128)])
def test_linear_dual_outputs(M, N, K, dt):
    """mumpy_linear_dual: shortcut sum (fp32) and the un-summed branch in operand precision from one pass."""
    ops = _ops()
    a = util.seeded_input((M, K), 1).to(dt)
    w = (util.seeded_input((N, K), 2) / K ** 0.5).to(dt)
    bias, res = util.seeded_input((N,), 3), util.seeded_input((M, N), 4)
    branch = orc.linear(a.float(), w.float(), bias)
    out, aux = ops.linear_dual(a.cuda(), w.cuda(), bias.cuda(), res.cuda())
    assert out.dtype == torch.float32 and aux.dtype == dt
    assert util.maxabs(out, branch + res) < 1e-4 * max(1.0, float(branch.abs().max()))
    assert util.maxabs(aux.float(), branch) < (1e-2 if dt == torch.bfloat16 else 2e-3) * max(1.0, float(branch.abs().max()))


LN_SHAPES = [  # (M, N, K): norm1 -> qkv and norm2 -> fc1 of the Swin blocks (stages 0-2 of the three views), ragged M
    (3136, 288, 96), (3136, 384, 96), (9408, 384, 128), (9408, 512, 128), (784, 576, 192), (784, 768, 192), (2352, 768, 256),
    (2352, 1024, 256), (196, 1152, 384), (6272, 1536, 384), (588, 1536, 512), (18816, 2048, 512), (130, 64, 128), (1, 192, 96),
]


@pytest.fixture
def ln_pair_mode(request):
    """CTA-pair policy of mumpy_ln_linear for one test (0 = 1-CTA tiles, 2 = cta_group::2 pairs whenever legal)."""
    ops = _ops()
    ops.set_ln_linear_pair_mode(request.param)
    yield request.param
    ops.set_ln_linear_pair_mode(1)


@pytest.mark.parametrize("ln_pair_mode", [0, 2], indirect=True)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", LN_SHAPES)
@pytest.mark.parametrize("gelu", [False, True])
def test_ln_linear_fused(M, N, K, gelu, dt, ln_pair_mode):
    """mumpy_ln_linear (LayerNorm produced in shared memory as the tcgen05 A operand) against (1) the oracle's fp32
    layer_norm -> round to operand precision -> linear (-> GELU) on the same seeded tensors, and (2) the unfused kernels
    (mumpy_layernorm + mumpy_linear), which it must reproduce bit for bit: same statistics arithmetic, same k-block order."""
    ops = _ops()
    x = util.seeded_input((M, K), 1) * 3.0 + util.seeded_input((M, 1), 5)
    g, b = 1.0 + 0.2 * util.seeded_input((K,), 6), 0.1 * util.seeded_input((K,), 7)
    w = (util.seeded_input((N, K), 2) / K ** 0.5).to(dt)
    bias = util.seeded_input((N,), 3)
    xn = orc.layer_norm(x, g, b).to(dt).float()
    ref = orc.linear(xn, w.float(), bias)
    if gelu:
        ref = orc.gelu(ref)
    act = ops.ACT_GELU if gelu else ops.ACT_NONE
    xc, gc, bc, wc, biasc = x.cuda(), g.cuda(), b.cuda(), w.cuda(), bias.cuda()
    out = ops.ln_linear(xc, gc, bc, 1e-5, wc, biasc, act=act)
    assert out.dtype == dt and tuple(out.shape) == (M, N)
    tol = 2e-2 if dt == torch.bfloat16 else 3e-3
    assert util.maxabs(out.float(), ref) < tol * max(1.0, float(ref.abs().max()))
    unfused = ops.linear(ops.layernorm(xc, gc, bc, 1e-5, out_dtype=dt), wc, biasc, act=act, out_dtype=dt)
    assert torch.equal(out, unfused)


@pytest.mark.parametrize("ln_pair_mode", [0, 2], indirect=True)
def test_ln_linear_many_items_per_cta(ln_pair_mode):
    """More work items than SMs (the persistent loop re-normalises into the same shared-memory operand) and no bias."""
    ops = _ops()
    M, N, K = 128 * 400 + 17, 384, 128
    x = util.seeded_input((M, K), 11).cuda()
    g, b = (1.0 + 0.1 * util.seeded_input((K,), 12)).cuda(), (0.1 * util.seeded_input((K,), 13)).cuda()
    w = (util.seeded_input((N, K), 14) / K ** 0.5).half().cuda()
    out = ops.ln_linear(x, g, b, 1e-5, w, None)
    unfused = ops.linear(ops.layernorm(x, g, b, 1e-5, out_dtype=torch.float16), w, None, out_dtype=torch.float16)
    assert torch.equal(out, unfused)


def test_ln_linear_rejects_unsupported_shapes():
    import mumpy_b200
    ops = _ops()
    x = torch.zeros((64, 768)).cuda()
    with pytest.raises(mumpy_b200._lib.MumpyError):
        ops.ln_linear(x, torch.ones(768).cuda(), torch.zeros(768).cuda(), 1e-5, torch.zeros((768, 768), dtype=torch.float16).cuda())
    was, was_w = ops.FUSED_LN, ops.FUSED_LN_WIDTHS
    try:
        ops.set_fused_ln(True, widths=(96, 128, 192, 256, 384, 512))
        assert not ops.ln_linear_fits(768, 768) and ops.ln_linear_fits(1536, 512) and ops.ln_linear_fits(384, 128)
        ops.set_fused_ln(True, widths=(512,))                      # the default: only the K = 512 blocks
        assert ops.ln_linear_fits(1536, 512) and not ops.ln_linear_fits(384, 128)
        ops.set_fused_ln(False)
        assert not ops.ln_linear_fits(1536, 512)
    finally:
        ops.set_fused_ln(was, widths=was_w)


def test_fused_ln_forward_is_bit_identical():
    """A Swin block and a CrossSwin block with the LayerNorms fused into qkv / fc1 (mumpy_ln_linear) equal the unfused kernels bit
    for bit (same statistics arithmetic, same k-block order)."""
    from mumpy_b200.models.modules.swinTransformer import SwinTransformerBlock
    ops = _ops()
    blk = SwinTransformerBlock(128, (14, 14), 4, window_size=7, shift_size=3, temporal_dim=3).eval()
    util.load_seeded(blk)
    blk = blk.cuda()
    x = util.seeded_input((2, 3 * 14 * 14, 128), 7).cuda()
    was, was_w = ops.FUSED_LN, ops.FUSED_LN_WIDTHS
    try:
        with torch.no_grad():
            ops.set_fused_ln(False)
            ref = blk(x).clone()
            ops.set_fused_ln(True, widths=(128,))
            out = blk(x)
        assert torch.equal(out, ref)
    finally:
        ops.set_fused_ln(was, widths=was_w)


@pytest.fixture
def mlp_shape(request):
    """Kernel shape of mumpy_mlp_fused for one test: 0 pipelined (default), 1 serial, 2 serial with two CTAs per SM."""
    ops = _ops()
    ops.set_mlp_fused_shape(request.param)
    yield request.param
    ops.set_mlp_fused_shape(0)


@pytest.mark.parametrize("mlp_shape", [0, 1, 2], indirect=True)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,C", [(3136, 96), (9408 + 17, 128), (784, 192), (2352, 256), (130, 128), (1, 96), (128 * 300 + 5, 128), (128 * 160, 256)])
def test_mlp_fused(M, C, dt, mlp_shape):
    """mumpy_mlp_fused (LayerNorm -> fc1 -> GELU -> fc2 -> + x in one kernel, hidden activations in shared / tensor memory only)
    against (1) the oracle's fp32 arithmetic on operands rounded where the kernel rounds them and (2) the three unfused kernels,
    which it must reproduce bit for bit (same statistics arithmetic, same k-block order in both GEMMs)."""
    ops = _ops()
    x = util.seeded_input((M, C), 1) * 2.0 + util.seeded_input((M, 1), 5)
    g, b = 1.0 + 0.2 * util.seeded_input((C,), 6), 0.1 * util.seeded_input((C,), 7)
    w1 = (util.seeded_input((4 * C, C), 2) / C ** 0.5).to(dt)
    w2 = (util.seeded_input((C, 4 * C), 3) / (4 * C) ** 0.5).to(dt)
    b1, b2 = util.seeded_input((4 * C,), 8), util.seeded_input((C,), 9)
    xn = orc.layer_norm(x, g, b).to(dt).float()
    h = orc.gelu(orc.linear(xn, w1.float(), b1)).to(dt).float()
    ref = x + orc.linear(h, w2.float(), b2)
    xc, gc, bc, w1c, w2c, b1c, b2c = x.cuda(), g.cuda(), b.cuda(), w1.cuda(), w2.cuda(), b1.cuda(), b2.cuda()
    out = ops.mlp_fused(xc, gc, bc, 1e-5, w1c, b1c, w2c, b2c)
    assert out.dtype == torch.float32 and tuple(out.shape) == (M, C)
    tol = 3e-2 if dt == torch.bfloat16 else 4e-3          # one operand-precision rounding of h feeds a K = 4C reduction
    assert util.maxabs(out, ref) < tol * max(1.0, float(ref.abs().max()))
    hid = ops.linear(ops.layernorm(xc, gc, bc, 1e-5, out_dtype=dt), w1c, b1c, act=ops.ACT_GELU, out_dtype=dt)
    unfused = ops.linear(hid, w2c, b2c, residual=xc)
    assert torch.equal(out, unfused)


def test_fused_mlp_block_is_bit_identical():
    """A stage-0 Swin block with the fused MLP kernel equals the three-kernel path bit for bit."""
    from mumpy_b200.models.modules.swinTransformer import SwinTransformerBlock
    ops = _ops()
    blk = SwinTransformerBlock(128, (14, 14), 4, window_size=7, shift_size=3, temporal_dim=3).eval()
    util.load_seeded(blk)
    blk = blk.cuda()
    x = util.seeded_input((2, 3 * 14 * 14, 128), 7).cuda()
    was = ops.FUSED_MLP
    try:
        with torch.no_grad():
            ops.set_fused_mlp(False)
            ref = blk(x).clone()
            ops.set_fused_mlp(True)
            out = blk(x)
        assert torch.equal(out, ref)
    finally:
        ops.set_fused_mlp(was)
