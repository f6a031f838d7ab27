"""GPU: each mirror module through the C ABI (fp32 mode) against the fixtures captured from the reference,
and against the oracle on the same seeded inputs.  Bar: max-abs <= 1e-4 (fp32 parity mode, BASELINE north_star)."""
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def mods():
    return util.golden("modules.pt")


@pytest.fixture(autouse=True)
def fp32_mode():
    import mumpy_b200
    mumpy_b200.set_precision("fp32")
    yield
    mumpy_b200.set_precision(mumpy_b200.ops.DEFAULT_PRECISION)


def _build(cls, ctor):
    m = cls(**ctor).eval()
    util.load_seeded(m)
    return m.cuda()


@pytest.mark.parametrize("size", [56, 224])
def test_faf(mods, size):
    from mumpy_b200.models.modules.dct import FAF
    f = mods["faf_%d" % size]
    x = util.seeded_input(f["input_shape"], f["input_seed"])
    y = FAF(size).eval().frame(x.cuda(), 1)
    if size == 224:
        assert util.maxabs(y[:, :, ::4, ::4], f["outputs"]["y_sub"]) < TOL
    else:
        assert util.maxabs(y, f["outputs"]["y"]) < TOL
    assert util.maxabs(y, orc.faf_middle(x)) < TOL


@pytest.mark.parametrize("name", ["swin_s0", "swin_s3", "swin_t1_s3", "swin_res7"])
def test_swin_block(mods, name):
    from mumpy_b200.models.modules.swinTransformer import SwinTransformerBlock
    f = mods[name]
    m = _build(SwinTransformerBlock, f["ctor"])
    x = util.seeded_input(f["input_shape"], f["input_seed"])
    assert util.maxabs(m(x.cuda()), f["outputs"]["y"]) < TOL


@pytest.mark.parametrize("name", ["sda_r3", "sda_r1"])
def test_swin_dattention(mods, name):
    from mumpy_b200.models.modules.deformableAttention import SwinDAttention
    f = mods[name]
    m = _build(SwinDAttention, f["ctor"])
    x1 = util.seeded_input(f["x1_shape"], f["input_seed"])
    x2 = util.seeded_input(f["x2_shape"], f["input_seed"] + 100)
    y, _ = m(x1.cuda(), x2.cuda())
    assert util.maxabs(y, f["outputs"]["y"]) < TOL


@pytest.mark.parametrize("name", ["cross_r3", "cross_r1", "cross_last"])
def test_cross_swin_block(mods, name):
    from mumpy_b200.models.encoder.multiTemporalViewEncoder import CrossSwinBlock
    f = mods[name]
    m = _build(CrossSwinBlock, f["ctor"])
    x1 = util.seeded_input(f["x1_shape"], f["input_seed"])
    x2 = x1 if f["ctor"]["last_view"] else util.seeded_input(f["x2_shape"], f["input_seed"] + 100)
    y, out = m(x1.cuda(), x2.cuda())
    assert util.maxabs(y, f["outputs"]["y"]) < TOL
    assert util.maxabs(out, f["outputs"]["out"]) < TOL


def test_patch_merging(mods):
    from mumpy_b200.models.modules.swinTransformer import PatchMerging
    f = mods["merge"]
    m = _build(PatchMerging, f["ctor"])
    x = util.seeded_input(f["input_shape"], f["input_seed"])
    assert util.maxabs(m(x.cuda()), f["outputs"]["y"]) < TOL


def test_vit_block(mods):
    from mumpy_b200.models.modules.blocks import Block
    f = mods["vit_block"]
    m = _build(Block, f["ctor"])
    x = util.seeded_input(f["input_shape"], f["input_seed"])
    assert util.maxabs(m(x.cuda()), f["outputs"]["y"]) < TOL


@pytest.fixture(scope="module")
def attn_maps():
    return util.golden("attn_maps.pt")


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("fp16", 3e-3)])
@pytest.mark.parametrize("name", ["sda_r3", "sda_r1"])
def test_swin_dattention_attention_map(attn_maps, name, mode, tol):
    """SwinDAttention.forward's second result (deformableAttention.py:399,405) against the map captured from the reference."""
    import mumpy_b200
    from mumpy_b200.models.modules.deformableAttention import SwinDAttention
    f = attn_maps[name]
    mumpy_b200.set_precision(mode)
    m = _build(SwinDAttention, f["ctor"])
    x1 = util.seeded_input(f["x1_shape"], f["input_seed"])
    x2 = util.seeded_input(f["x2_shape"], f["input_seed"] + 100)
    _, attn = m(x1.cuda(), x2.cuda())
    assert attn.dtype == torch.float32 and tuple(attn.shape) == tuple(f["attn"].shape)
    assert util.maxabs(attn, f["attn"]) < tol
    assert util.maxabs(attn.sum(-1), torch.ones(attn.shape[:-1])) < 1e-5


def test_cva_module_return_attention(attn_maps):
    """CVAModule.forward(return_attention=True) returns the map alone (multiTemporalViewEncoder.py:134-137)."""
    from mumpy_b200.models.encoder.multiTemporalViewEncoder import CVAModule
    f = attn_maps["cva_module"]
    m = _build(CVAModule, f["ctor"])
    x1 = util.seeded_input(f["x1_shape"], f["input_seed"])
    x2 = util.seeded_input(f["x2_shape"], f["input_seed"] + 100)
    attn = m(x1.cuda(), x2.cuda(), return_attention=True)
    assert util.maxabs(attn, f["attn"]) < 2e-5
    y, attn2 = m(x1.cuda(), x2.cuda())
    assert torch.equal(attn, attn2) and tuple(y.shape) == tuple(x1.shape)


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("fp16", 3e-3)])
def test_vit_block_return_attention(attn_maps, mode, tol):
    """Block.forward(return_attention=True) (blocks.py:86-89) and Attention.forward's second result (:66-74)."""
    import mumpy_b200
    from mumpy_b200.models.modules.blocks import Block
    f = attn_maps["vit_block"]
    mumpy_b200.set_precision(mode)
    m = _build(Block, f["ctor"])
    x = util.seeded_input(f["input_shape"], f["input_seed"]).cuda()
    attn = m(x, return_attention=True)
    assert attn.dtype == torch.float32 and tuple(attn.shape) == tuple(f["attn"].shape)
    assert util.maxabs(attn, f["attn"]) < tol
    xn = torch.nn.functional.layer_norm(x, (x.shape[-1],), m.norm1.weight, m.norm1.bias, m.norm1.eps)
    y, attn2 = m.attn(xn)
    assert tuple(y.shape) == tuple(x.shape) and util.maxabs(attn2, f["attn"]) < max(tol, 1e-4)


def test_tokenize(mods):
    from mumpy_b200.models.encoder.multiTemporalViewEncoder import CrossThreeViewTokenize
    from mumpy_b200.models.factory.modelFactory import default_view_configs
    f = mods["tokenize"]
    m = CrossThreeViewTokenize(default_view_configs()).eval()
    util.load_seeded(m)
    m = m.cuda()
    x = util.seeded_input(f["input_shape"], f["input_seed"])
    out = m(x.cuda())
    for i in range(3):
        assert util.maxabs(out[i].reshape(2, -1, out[i].shape[-1]), f["outputs"]["v%d" % i]) < TOL


@pytest.mark.parametrize("mode,tol", [("fp16", 2e-5), ("bf16", 3e-4)])
def test_tokenize_tensor_core_path(mods, mode, tol):
    """16-bit modes: the tokenizer convolution as split-operand tcgen05 GEMMs (patches [hi|hi|lo] x weights [hi|lo|hi], ~22 / ~16
    mantissa bits) + fp32 LayerNorm, against the reference capture; and the fp32 FMA kernel it replaces gives the same tokens."""
    import mumpy_b200
    from mumpy_b200.models.encoder import multiTemporalViewEncoder as mtv
    from mumpy_b200.models.factory.modelFactory import default_view_configs
    f = mods["tokenize"]
    m = mtv.CrossThreeViewTokenize(default_view_configs()).eval()
    util.load_seeded(m)
    m = m.cuda()
    x = util.seeded_input(f["input_shape"], f["input_seed"]).cuda()
    mumpy_b200.set_precision(mode)
    try:
        out = m(x)
        mtv.TENSOR_CORE_TOKENIZER = False
        fma = m(x)
    finally:
        mtv.TENSOR_CORE_TOKENIZER = True
        mumpy_b200.set_precision(mumpy_b200.ops.DEFAULT_PRECISION)
    for i in range(3):
        assert out[i].dtype == torch.float32 and out[i].shape == fma[i].shape
        assert util.maxabs(out[i].reshape(2, -1, out[i].shape[-1]), f["outputs"]["v%d" % i]) < tol
        assert util.maxabs(out[i], fma[i]) < tol


def test_decoder_feats_view_equals_transposed_copy():
    """Decoder.forward hands x_feats back as a channels-last view with the reference's shape; NCHW_FEATS = True restores the
    explicit transposition: identical values."""
    import mumpy_b200
    from mumpy_b200.models.decoder.decoder import Decoder
    dec = Decoder().eval()
    util.load_seeded(dec)
    dec = dec.cuda()
    B = 1
    final_x, view_x, ff = util.decoder_inputs(B)
    x, view_x, ff = final_x.cuda(), [[t.cuda() for t in st] for st in view_x], ff.cuda()
    with torch.no_grad():
        m0, f0 = dec(x, view_x, ff)
        Decoder.NCHW_FEATS = True
        try:
            m1, f1 = dec(x, view_x, ff)
        finally:
            Decoder.NCHW_FEATS = False
    assert f0.shape == f1.shape == (B, 32, 224, 224) and f1.is_contiguous() and not f0.is_contiguous()
    assert torch.equal(f0, f1) and torch.equal(m0, m1)


def test_window_attention_with_mask_api():
    """WindowAttention.forward(x_windows, mask) keeps the reference call form (swinTransformer.py:134-166)."""
    from mumpy_b200.models.modules.swinTransformer import WindowAttention
    m = WindowAttention(64, (7, 7), 2).eval()
    sd = util.load_seeded(m)
    m = m.cuda()
    x = util.seeded_input((8, 49, 64), 5)
    mask = orc.shifted_window_mask(14, 14, 7, 3)
    ref = orc.window_attention(sd, "", x, 2, 7, mask)
    assert util.maxabs(m(x.cuda(), mask.cuda()), ref) < TOL


def test_decoder(mods):
    from mumpy_b200 import Decoder
    f = mods["decoder"]
    m = Decoder().eval()
    util.load_seeded(m)
    m = m.cuda()
    final_x, view_x, ff = util.decoder_inputs(1)
    lg, xf = m(final_x.cuda(), [[t.cuda() for t in st] for st in view_x], ff.cuda())
    assert util.maxabs(lg, f["outputs"]["logits"]) < TOL
    assert util.maxabs(xf[:, :, ::4, ::4], f["outputs"]["x_feats_sub"]) < TOL
