"""GPU: the path at BASELINE.json's full sizes (configs[2]: batch 32 at 224x224; configs[1]: isolated kernels at batch 64),
where the CPU oracle would take minutes, checked through size-independent properties of the domain:

  * a clip's result does not depend on the batch it is processed in (clips are independent in test.py; with per-clip window
    pairing in the deformable attention -- SURVEY 8(e) caveat -- this holds for the whole forward), so rows of a batch-32
    forward must equal batch-1 forwards of the same clips, which the golden-vector tests pin to the reference;
  * softmax rows sum to one: window attention with V = 1 returns 1, for plain and shifted (masked) windows;
  * GEMM / DCT kernels commute with scaling by powers of two and with row permutations, bit for bit;
  * the resize of a constant image is that constant, and one frame of a full batch equals the oracle's;
  * the integer mask counts satisfy n_union = n_pred + n_gt - TP.
"""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    import mumpy_b200
    enc, dec = mumpy_b200.Encoder().eval(), mumpy_b200.Decoder().eval()
    util.load_seeded(enc)
    util.load_seeded(dec)
    return enc.cuda(), dec.cuda()


@pytest.mark.parametrize("mode,tol,identity", [("fp32", 2e-5, 0.9999), ("fp16", 4e-3, 0.999), ("bf16", 3e-2, 0.994)])
def test_batch32_rows_equal_single_clip_forwards(model, mode, tol, identity):
    """configs[2] shape.  fp32 mode: the only batch dependence is fp32 summation order (split-K ranges, tile widths).  16-bit
    modes: tile choices change with M, rounding points do not, so the tolerances of the golden-vector tests apply with room."""
    import mumpy_b200
    from mumpy_b200.models.encoder import multiTemporalViewEncoder as mtv
    enc, dec = model
    B = 32
    x = util.seeded_input((B, 3, 3, 224, 224), 4321).cuda()
    mumpy_b200.set_precision(mode)
    mtv.set_per_clip_pairing(True)
    try:
        with torch.no_grad():
            full = mumpy_b200.forward(enc, dec, x)[0].clone()
            assert full.shape == (B, 1, 224, 224) and bool(torch.isfinite(full).all())
            for i in (0, 13, 31):
                one = mumpy_b200.forward(enc, dec, x[i:i + 1].contiguous())[0]
                assert util.maxabs(full[i:i + 1], one) < tol, (mode, i)
                same = ((full[i:i + 1] > 0) == (one > 0)).float().mean().item()
                assert same >= identity, (mode, i, same)
    finally:
        mtv.set_per_clip_pairing(False)
        mumpy_b200.set_precision(mumpy_b200.ops.DEFAULT_PRECISION)


def test_batch32_mask_counts_identity(model):
    import mumpy_b200
    enc, dec = model
    x = util.seeded_input((32, 3, 3, 224, 224), 77).cuda()
    gt = (util.seeded_input((32, 224, 224), 78) > 0.2).to(torch.uint8).cuda()
    with torch.no_grad():
        logits = mumpy_b200.forward(enc, dec, x)[0]
    mask, counts = mumpy_b200.ops.mask_counts(logits, gt)
    c = counts.cpu()
    assert torch.equal(c[:, 3], c[:, 1] + c[:, 2] - c[:, 0])                       # union = pred + gt - TP
    assert torch.equal(c[:, 1], (mask.cpu() > 0).flatten(1).sum(1))
    assert torch.equal(c[:, 2], (gt.cpu() > 0).flatten(1).sum(1))
    assert bool(((mask == 0) | (mask == 255)).all())


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shift", [0, 3])
def test_window_attention_b64_rows_sum_to_one(dt, shift):
    """configs[1] (iii): stage 0 of view 3 at batch 64 (12288 windows x 4 heads), V = 1 -> every output is sum_j P_ij / sum_j P_ij.
    P is rounded to the operand type before PV while the normaliser uses the unrounded values: |out - 1| <= 2^-8 (bf16) / 2^-11 (f16)."""
    from mumpy_b200 import ops
    B, TH, W, C, heads = 64, 168, 56, 128, 4
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn((B, TH * W, 3 * C), device="cuda", generator=g).to(dt)
    qkv[:, :, 2 * C:] = 1.0
    table = (0.5 * torch.randn((169, heads), device="cuda", generator=g)).contiguous()
    bias = torch.zeros((heads, 49, 49), device="cuda")
    mask = torch.zeros(((TH // 7) * (W // 7), 49, 49), device="cuda") if shift else None
    out = ops.window_attention(qkv, bias, mask, B, TH, W, C, heads, 7, shift, rel_table=table, standard_mask=shift > 0)
    assert out.shape == (B, TH * W, C)
    assert float((out.float() - 1.0).abs().max()) <= (2 ** -8 if dt == torch.bfloat16 else 2 ** -11)


@pytest.mark.parametrize("M,N,K,epi", [(602112, 384, 128, "plain"), (602112, 128, 512, "residual"), (37632, 2048, 512, "gelu16")])
def test_linear_b64_scaling_and_row_permutation(M, N, K, epi):
    """configs[1] (iv) GEMM shapes at batch 64.  Scaling the A operand by 2 scales the (bias-free, activation-free) product by
    exactly 2, and permuting the rows of A permutes the rows of the output -- bit for bit, whatever tile a row lands in."""
    from mumpy_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    w = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5).bfloat16()
    perm = torch.randperm(M, device="cuda", generator=g)
    if epi == "gelu16":
        f = lambda t: ops.linear(t, w, None, act=ops.ACT_GELU, out_dtype=torch.bfloat16)
    elif epi == "residual":
        res = torch.zeros((M, N), device="cuda")
        f = lambda t: ops.linear(t, w, None, residual=res)
    else:
        f = lambda t: ops.linear(t, w, None)
    y = f(a)
    assert torch.equal(f(a[perm].contiguous()), y[perm])
    if epi != "gelu16":
        assert torch.equal(f((a * 2).contiguous()), y * 2)
    assert bool(torch.isfinite(y.float()).all())


def test_faf_b32_scaling_and_frame_independence():
    """DCT branch at batch 32: exact under scaling by two (fp32 and hi/lo-split arithmetic alike), and a clip's band maps do not
    depend on its neighbours in the batch."""
    import mumpy_b200
    from mumpy_b200.models.modules.dct import FAF
    faf = FAF(224).eval()
    x = util.seeded_input((32, 3, 3, 224, 224), 31).cuda()
    for mode in ("bf16", "fp32"):
        mumpy_b200.set_precision(mode)
        try:
            with torch.no_grad():
                y = faf.frame(x, 1).clone()
                assert y.shape == (32, 9, 224, 224)
                assert torch.equal(faf.frame((x * 2).contiguous(), 1), y * 2)
                assert torch.equal(faf.frame(x[5:6].contiguous(), 1), y[5:6])
        finally:
            mumpy_b200.set_precision(mumpy_b200.ops.DEFAULT_PRECISION)


def test_resize_b64_native_frames():
    """64 native-resolution DAVIS frames (480 x 854): constants stay constant, one random frame equals the oracle."""
    from mumpy_b200 import ops
    from oracle import pil_resample as pr
    frames = torch.empty((64, 480, 854, 3), dtype=torch.uint8)
    for i in range(64):
        frames[i] = (i * 4) % 256
    img = pr.seeded_image(64, 480, 854, 3)
    frames[37] = torch.from_numpy(img)
    for code, fn in ((ops.RESIZE_BICUBIC, pr.resize_bicubic_u8), (ops.RESIZE_NEAREST, pr.resize_nearest_u8)):
        out = ops.resize_u8(frames.cuda(), 224, 224, code).cpu()
        for i in (0, 1, 36, 38, 63):
            assert bool((out[i] == (i * 4) % 256).all()), i
        assert np.array_equal(out[37].numpy(), fn(img, 224, 224))
