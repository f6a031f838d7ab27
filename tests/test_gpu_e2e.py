"""GPU: the full forward (Encoder -> Decoder, the path test.py:94-95 drives) through the C ABI against the reference
captures.  fp32 mode: <= 1e-4 on logits and identical masks.  fp16 mode (the default and the mode bench.py reports):
max-abs <= 5e-3, mean-abs <= 6e-4, >= 99.9 % identical thresholded masks (the north star's bar).  bf16 mode (opt-in):
its own stated, weaker tolerance."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

# bf16-mode tolerances (tcgen05 GEMMs on bf16 operands, fp32 accumulation / residual stream / statistics),
# measured on B200 with the key-seeded weights, logit std 0.18:
# bf16 is the opt-in wide-range mode; it does NOT meet the north star's >= 99.9 % mask identity (8 mantissa bits on every
# operand: measured 1.0e-2 / 1.3e-2 max-abs, 1.9e-3 mean-abs, 99.60 % identity) -- the default / benchmarked mode is fp16 (below).
BF16_LOGIT_MAXABS = 3e-2
BF16_LOGIT_MEANABS = 3e-3
BF16_MASK_IDENTITY = 0.994


@pytest.fixture(scope="module")
def model():
    import mumpy_b200
    enc, dec = mumpy_b200.Encoder().eval(), mumpy_b200.Decoder().eval()
    util.load_seeded(enc)
    util.load_seeded(dec)
    return enc.cuda(), dec.cuda()


def _run(model, x, mode):
    import mumpy_b200
    mumpy_b200.set_precision(mode)
    try:
        enc, dec = model
        with torch.no_grad():
            final_x, view_x, ff = enc(x.cuda())
            logits, feats = dec(final_x, view_x, ff)
        torch.cuda.synchronize()
        return final_x, view_x, ff, logits, feats
    finally:
        mumpy_b200.set_precision(mumpy_b200.ops.DEFAULT_PRECISION)


@pytest.mark.parametrize("B", [1, 2])
def test_fp32_mode_matches_reference(model, B):
    g = util.golden("e2e_b%d.pt" % B)
    x = util.seeded_input(g["input_shape"], g["input_seed"])
    final_x, view_x, ff, logits, feats = _run(model, x, "fp32")
    assert final_x.shape == (B, 2304, 7, 7) and logits.shape == (B, 1, 224, 224) and feats.shape == (B, 32, 224, 224)
    assert util.maxabs(ff[:, :, ::4, ::4], g["ffinfo_sub"]) < 1e-4
    for s in range(4):
        for v in range(3):
            assert view_x[s][v].shape[:2] == (B, 1)
            assert util.maxabs(view_x[s][v][:, :, ::16, :], g["view_sub"][s][v]) < 5e-4, (s, v)
    assert util.maxabs(final_x, g["final_x"]) < 5e-4
    assert util.maxabs(logits, g["logits"]) < 1e-4
    assert util.maxabs(feats[:, :, ::4, ::4], g["x_feats_sub"]) < 1e-4
    assert int(((logits.cpu() > 0) != (g["logits"] > 0)).sum()) <= int(1e-3 * logits.numel())


@pytest.mark.parametrize("B", [1, 2])
def test_bf16_mode_within_stated_tolerance(model, B):
    g = util.golden("e2e_b%d.pt" % B)
    x = util.seeded_input(g["input_shape"], g["input_seed"])
    _, _, _, logits, _ = _run(model, x, "bf16")
    d = (logits.cpu() - g["logits"]).abs()
    same = float(((logits.cpu() > 0) == (g["logits"] > 0)).float().mean())
    print("bf16 B=%d: logits max-abs %.3e mean-abs %.3e mask identity %.5f" % (B, float(d.max()), float(d.mean()), same))
    assert float(d.max()) < BF16_LOGIT_MAXABS
    assert float(d.mean()) < BF16_LOGIT_MEANABS
    assert same > BF16_MASK_IDENTITY


# fp16 mode (IEEE-half operands, same kernels and speed): 3 more mantissa bits than bf16 on every GEMM operand.
# Measured on B200 (round 1): max-abs 1.1e-3 / 1.3e-3, mean-abs 2.1e-4, mask identity 99.954 % / 99.951 % (B=1 / B=2)
# -- the north star's >= 99.9 % pixel-identity bar, on logits deliberately centred on the threshold.
FP16_LOGIT_MAXABS = 5e-3
FP16_LOGIT_MEANABS = 6e-4
FP16_MASK_IDENTITY = 0.999


@pytest.mark.parametrize("B", [1, 2])
def test_fp16_mode_within_stated_tolerance(model, B):
    g = util.golden("e2e_b%d.pt" % B)
    x = util.seeded_input(g["input_shape"], g["input_seed"])
    _, _, _, logits, _ = _run(model, x, "fp16")
    d = (logits.cpu() - g["logits"]).abs()
    same = float(((logits.cpu() > 0) == (g["logits"] > 0)).float().mean())
    print("fp16 B=%d: logits max-abs %.3e mean-abs %.3e mask identity %.5f" % (B, float(d.max()), float(d.mean()), same))
    assert float(d.max()) < FP16_LOGIT_MAXABS
    assert float(d.mean()) < FP16_LOGIT_MEANABS
    assert same > FP16_MASK_IDENTITY


@pytest.mark.parametrize("name", ["e2e_256w8_b1", "e2e_512w8_b1"])
def test_patched_resolution_window8(name):
    """BASELINE.json configs[3] / SURVEY A10: window 8 at 256x256 and 512x512 (DVI resolution) against captures of the
    reference classes instantiated with the same patched configuration (oracle/make_golden_hires.py).
    fp32 mode <= 1e-4 on logits and identical masks; the 16-bit modes within their stated tolerances."""
    import mumpy_b200
    g = util.golden(name + ".pt")
    size, ws, res = g["size"], g["window"], g["res"]
    enc = mumpy_b200.Encoder(img_size=size, window_size=ws).eval()
    dec = mumpy_b200.Decoder(shape=res).eval()
    util.load_seeded(enc.base)            # the fixture seeds the bare ThreeViewSwinTransformer (keys without "base.")
    util.load_seeded(dec)
    m = (enc.cuda(), dec.cuda())
    x = util.seeded_input(g["input_shape"], g["input_seed"])
    final_x, view_x, ff, logits, feats = _run(m, x, "fp32")
    assert logits.shape == (1, 1, size, size) and final_x.shape == (1, 2304, res[3], res[3])
    assert util.maxabs(ff[:, :, ::8, ::8], g["ffinfo_sub"]) < 1e-4
    for s in range(4):
        for v in range(3):
            assert util.maxabs(view_x[s][v][:, :, ::64, :], g["view_sub"][s][v]) < 5e-4, (s, v)
    assert util.maxabs(final_x, g["final_x"]) < 5e-4
    assert util.maxabs(logits, g["logits"]) < 1e-4
    assert util.maxabs(feats[:, :, ::8, ::8], g["x_feats_sub"]) < 1e-4
    assert int(((logits.cpu() > 0) != (g["logits"] > 0)).sum()) <= int(1e-3 * logits.numel())
    for mode, tol_max, tol_mean, ident in (("fp16", FP16_LOGIT_MAXABS, FP16_LOGIT_MEANABS, FP16_MASK_IDENTITY),
                                           ("bf16", BF16_LOGIT_MAXABS, BF16_LOGIT_MEANABS, BF16_MASK_IDENTITY)):
        _, _, _, lg, _ = _run(m, x, mode)
        d = (lg.cpu() - g["logits"]).abs()
        same = float(((lg.cpu() > 0) == (g["logits"] > 0)).float().mean())
        print("%s %s: logits max-abs %.3e mean-abs %.3e mask identity %.5f" % (name, mode, float(d.max()), float(d.mean()), same))
        assert float(d.max()) < tol_max and float(d.mean()) < tol_mean and same > ident


def test_mask_and_counts(model):
    """a20 + measure.py:77-91: thresholded mask and integer counts from the fused kernel vs the oracle."""
    import mumpy_b200
    from oracle import mumpy_oracle as orc
    g = util.golden("e2e_b2.pt")
    logits = g["logits"]
    gt = (util.seeded_input((2, 224, 224), 99) > 0.3)
    mask, counts = mumpy_b200.ops.mask_counts(logits.cuda(), gt.to(torch.uint8).cuda())
    assert torch.equal(mask.cpu(), orc.threshold_mask(logits)[:, 0])
    assert torch.equal(counts.cpu(), orc.clip_counts(logits[:, 0] > 0, gt))


def test_graph_replay_and_lanes_are_deterministic(model):
    """The forward captured in a CUDA graph (fork/join lanes become parallel branches, as bench.py runs it) replays to
    exactly the eager result, the eager result is reproducible run to run, and the single-stream schedule
    (streams.set_enabled(False)) gives the same bits: every kernel is deterministic and the lanes only reorder
    independent work."""
    import mumpy_b200
    from mumpy_b200 import streams
    enc, dec = model
    g = util.golden("e2e_b2.pt")
    x = util.seeded_input(g["input_shape"], g["input_seed"]).cuda()
    mumpy_b200.set_precision("bf16")

    def fwd():
        final_x, view_x, ff = enc(x)
        return dec(final_x, view_x, ff)[0]

    with torch.no_grad():
        a = fwd().clone()
        b = fwd().clone()
        assert torch.equal(a, b)
        streams.set_enabled(False)
        try:
            c = fwd().clone()
        finally:
            streams.set_enabled(True)
        assert torch.equal(a, c)
        # one outer region around both modules (decoder branches overlap the encoder's tail): same bits
        d2 = mumpy_b200.forward(enc, dec, x)[0].clone()
        assert torch.equal(a, d2)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fwd()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = mumpy_b200.forward(enc, dec, x)[0]
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, a)
