"""GPU: parity of the BENCHMARKED configuration and of the evaluation loop around it.

  * bench.py's own first batch (32 clips, seed 1234) with the reference's batch-global deformable pairing, compared with the
    oracle run at the same batch size on the host (~15 s): fp32 mode <= 1e-4, the headline mode (fp16) >= 99.9 % identical masks;
  * the fp16 range guard: leaving the half range saturates AND is reported (never silent), bf16 never reports;
  * mumpy_b200.forward on non-contiguous / half-precision inputs (the conversion must precede the lane fork);
  * ShardedEvaluator with the real model: per-clip counts against the oracle, identical for any micro-batch / shard split.
"""
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    import mumpy_b200
    enc, dec = mumpy_b200.Encoder().eval(), mumpy_b200.Decoder().eval()
    enc_sd, dec_sd = util.load_seeded(enc), util.load_seeded(dec)
    return enc.cuda(), dec.cuda(), enc_sd, dec_sd


def test_bench_batch32_global_pairing_vs_oracle(model):
    """configs[2] exactly as bench.py runs it: batch 32, PER_CLIP_PAIRING False (a clip's deformable branch sees windows of other
    clips of the batch, deformableAttention.py:329-330,394-395), bench.py's first seeded batch."""
    import bench
    import mumpy_b200
    enc, dec, enc_sd, dec_sd = model
    x = bench.synthetic_batches(32, 224, rank=0, n=1)[0]
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        ref, _ = orc.forward(enc_sd, dec_sd, x)                      # batch-global pairing is the oracle's default
        assert ref.shape == (32, 1, 224, 224)
        for mode, tol_max, tol_mean, ident in (("fp32", 1e-4, 2e-5, 0.9999), ("fp16", 6e-3, 6e-4, 0.999)):
            mumpy_b200.set_precision(mode)
            logits = mumpy_b200.forward(enc, dec, x.cuda())[0].cpu()
            mumpy_b200.ops.check_f16_range()
            d = (logits - ref).abs()
            same = float(((logits > 0) == (ref > 0)).float().mean())
            print("bench batch 32 %s vs oracle: max-abs %.3e mean-abs %.3e mask identity %.5f" % (mode, float(d.max()), float(d.mean()), same))
            assert float(d.max()) < tol_max and float(d.mean()) < tol_mean and same >= ident, mode
            per_clip = ((logits > 0) == (ref > 0)).float().flatten(1).mean(1)
            assert float(per_clip.min()) >= (0.9995 if mode == "fp32" else 0.997), (mode, float(per_clip.min()))


def test_fp16_overflow_saturates_and_is_reported():
    import mumpy_b200
    from mumpy_b200 import ops
    ops.f16_overflowed()                                           # clear
    big = torch.full((4096,), 1.0e5).cuda()
    y = ops.cast16(big, torch.float16)
    assert bool((y.float() == 65504.0).all())                      # saturated, not inf
    assert ops.f16_overflowed() is True
    assert ops.f16_overflowed() is False                           # reading resets
    # GEMM epilogue: activations x64 leave the range
    a = (util.seeded_input((256, 512), 1) * 40).half().cuda()
    w = util.seeded_input((1536, 512), 2).half().cuda()
    out = ops.linear(a, w, out_dtype=torch.float16)
    assert bool(torch.isfinite(out.float()).all()) and not ops.f16_overflowed()
    out = ops.linear((a * 64).contiguous(), w, out_dtype=torch.float16)
    assert bool(torch.isfinite(out.float()).all()) and float(out.float().abs().max()) == 65504.0
    with pytest.raises(mumpy_b200._lib.MumpyError):
        ops.check_f16_range()
    # LayerNorm output / gathers report too; bf16 never does
    ops.layernorm(torch.randn(64, 128).cuda(), torch.full((128,), 1.0e5).cuda(), torch.zeros(128).cuda(), out_dtype=torch.float16)
    assert ops.f16_overflowed()
    ops.cast16(big, torch.bfloat16)
    ops.linear((a.float() * 64).bfloat16(), w.bfloat16(), out_dtype=torch.bfloat16)
    assert not ops.f16_overflowed()


def test_sharded_evaluator_raises_on_fp16_overflow():
    import mumpy_b200
    from mumpy_b200 import evaluate as ev
    from mumpy_b200 import ops
    ops.f16_overflowed()
    big = torch.full((1024,), 3.0e5).cuda()

    def predict(clips):
        ops.cast16(big, torch.float16)                             # an activation beyond the half range somewhere in the forward
        return clips

    logits = torch.ones((4, 1, 8, 8)).cuda()
    gt = torch.ones((4, 8, 8), dtype=torch.uint8).cuda()
    e = ev.ShardedEvaluator(predict, lambda lo, hi: (logits[lo:hi], gt[lo:hi]), 4, 2)
    with pytest.raises(mumpy_b200._lib.MumpyError):
        e.run(device=torch.device("cuda"))


def test_forward_converts_inputs_before_the_lane_fork(model):
    """Non-contiguous and half-precision clip batches through mumpy_b200.forward (lanes fork inside) equal the contiguous fp32
    call, bit for bit, also when the conversion kernel is slow to start (a spin kernel parks the stream first)."""
    import mumpy_b200
    enc, dec = model[0], model[1]
    big = util.seeded_input((4, 3, 3, 224, 224), 5).cuda()
    with torch.no_grad():
        ref = mumpy_b200.forward(enc, dec, big[::2].contiguous())[0].clone()
        for _ in range(2):
            torch.cuda._sleep(int(2e7))
            out = mumpy_b200.forward(enc, dec, big[::2])[0]
            assert torch.equal(out, ref)
        xh = big[:2].half()
        refh = mumpy_b200.forward(enc, dec, xh.float().contiguous())[0].clone()
        torch.cuda._sleep(int(2e7))
        assert torch.equal(mumpy_b200.forward(enc, dec, xh)[0], refh)


def test_sharded_evaluator_real_model_vs_oracle(model):
    """configs[4] in the small: uint8 frames of two sequences -> ClipAssembler -> forward -> mask_counts through
    ShardedEvaluator, per-clip deformable pairing.  Counts equal the oracle's up to threshold-straddling pixels, and are
    IDENTICAL for every micro-batch size (= any shard split): the per-clip F1 / IoU do not depend on how clips are grouped."""
    import mumpy_b200
    from mumpy_b200 import evaluate as ev
    from mumpy_b200 import frontend
    from mumpy_b200.models.encoder import multiTemporalViewEncoder as mtv
    enc, dec, enc_sd, dec_sd = model
    seqs = [5, 3]
    g = torch.Generator().manual_seed(77)
    frames = torch.randint(0, 256, (sum(seqs), 224, 224, 3), generator=g, dtype=torch.uint8)
    gt = (torch.rand((sum(seqs), 224, 224), generator=g) > 0.6)
    asm = frontend.ClipAssembler(frames, seqs, torch.device("cuda"))
    gt_dev = gt.to(torch.uint8).cuda()
    mtv.set_per_clip_pairing(True)
    try:
        with torch.no_grad():
            clips = orc.assemble_clips(frames, seqs)
            ref_logits, _ = orc.forward(enc_sd, dec_sd, clips, per_clip_pairing=True)
            ref_counts = orc.clip_counts(ref_logits[:, 0] > 0, gt)
            tables = {}
            for mode in ("fp32", "fp16"):
                mumpy_b200.set_precision(mode)
                for mb in (8, 3, 1):
                    e = ev.ShardedEvaluator(lambda x: mumpy_b200.forward(enc, dec, x)[0], lambda lo, hi: (asm.batch(lo, hi), gt_dev[lo:hi]),
                                            len(asm), mb)
                    res = e.run(device=torch.device("cuda"))
                    tables[(mode, mb)] = res["counts"]
                    assert res["n_valid"] == 8
                # different micro-batch sizes: only fp32 summation order (tile widths follow M) differs -> threshold pixels at most
                for mb in (3, 1):
                    assert int((tables[(mode, 8)] - tables[(mode, mb)]).abs().max()) <= (2 if mode == "fp32" else 60), (mode, mb)
                # shard-aligned micro-batches (what bench.py's split run uses: 43 divides every shard): two "ranks" with two
                # micro-batches of 2 each reproduce the one-rank table bit for bit
                one = ev.ShardedEvaluator(lambda x: mumpy_b200.forward(enc, dec, x)[0], lambda lo, hi: (asm.batch(lo, hi), gt_dev[lo:hi]), 8, 2).run(device=torch.device("cuda"))["counts"]
                halves = [ev.ShardedEvaluator(lambda x: mumpy_b200.forward(enc, dec, x)[0],
                                              lambda lo, hi, o=o: (asm.batch(o + lo, o + hi), gt_dev[o + lo:o + hi]), 4, 2).run(device=torch.device("cuda"))["counts"]
                          for o in (0, 4)]
                assert torch.equal(one, torch.cat(halves, 0)), mode
                lim = 8 if mode == "fp32" else 60                                 # pixels of 50176 (0.1 % = 50)
                assert int((tables[(mode, 8)] - ref_counts).abs().max()) <= lim, (mode, tables[(mode, 8)], ref_counts)
            f1, iou = ev.f1_iou_from_counts(tables[("fp32", 8)], 224 * 224)
            f1o, iouo = orc.f1_iou_from_counts(ref_counts, 224 * 224)
            assert float((f1 - f1o).abs().max()) < 1e-3 and float((iou - iouo).abs().max()) < 1e-3
    finally:
        mtv.set_per_clip_pairing(False)
