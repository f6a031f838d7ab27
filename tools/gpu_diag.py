"""Development aid: runs every parity check on the GPU without stopping at the first failure and prints the
numbers (max-abs vs the reference captures / the oracle).  Usage: python tools/gpu_diag.py [--skip-bf16]"""
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mumpy_oracle as orc  # noqa: E402
from tests import util  # noqa: E402

import mumpy_b200  # noqa: E402
from mumpy_b200 import ops  # noqa: E402


def section(name):
    print("\n=== %s ===" % name, flush=True)


def guarded(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()
        return None


def gemm_checks(bf16):
    shapes = [(49, 96, 96), (3136, 288, 96), (3136, 96, 384), (9408, 384, 128), (588, 1536, 512), (588, 512, 2048),
              (147, 768, 2560), (130, 64, 72), (256, 128, 64), (128, 256, 128)]
    for M, N, K in shapes:
        a, w = util.seeded_input((M, K), 1), util.seeded_input((N, K), 2) / K ** 0.5
        bias, res = util.seeded_input((N,), 3), util.seeded_input((M, N), 4)
        if bf16:
            a, w = a.bfloat16(), w.bfloat16()
        ref = orc.gelu(orc.linear(a.float(), w.float(), bias)) + res
        try:
            out = ops.linear(a.cuda(), w.cuda(), bias.cuda(), res.cuda(), act=ops.ACT_GELU)
            torch.cuda.synchronize()
            err = (out.cpu() - ref).abs()
            print("M=%d N=%d K=%d %s: max-abs %.3e (ref max %.2f)" % (M, N, K, "bf16" if bf16 else "fp32", float(err.max()), float(ref.abs().max())), flush=True)
            if float(err.max()) > 1e-2:
                rb = err.reshape(M, N).max(1).values
                cb = err.reshape(M, N).max(0).values
                print("   bad rows (first 16 of %d): %s" % (int((rb > 1e-2).sum()), (rb > 1e-2).nonzero().flatten()[:16].tolist()))
                print("   bad cols (first 16 of %d): %s" % (int((cb > 1e-2).sum()), (cb > 1e-2).nonzero().flatten()[:16].tolist()))
                print("   out[0,:8]", out[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())
        except Exception:
            traceback.print_exc()
            return False
    return True


def module_checks():
    mods = util.golden("modules.pt")
    from mumpy_b200.models.modules.dct import FAF
    from mumpy_b200.models.modules.swinTransformer import SwinTransformerBlock, PatchMerging
    from mumpy_b200.models.modules.deformableAttention import SwinDAttention
    from mumpy_b200.models.modules.blocks import Block
    from mumpy_b200.models.encoder.multiTemporalViewEncoder import CrossSwinBlock, CrossThreeViewTokenize
    from mumpy_b200.models.factory.modelFactory import default_view_configs

    def build(cls, ctor):
        m = cls(**ctor).eval()
        util.load_seeded(m)
        return m.cuda()

    def faf(size):
        f = mods["faf_%d" % size]
        x = util.seeded_input(f["input_shape"], f["input_seed"])
        y = FAF(size).eval().frame(x.cuda(), 1)
        print("faf_%d vs oracle: %.3e" % (size, util.maxabs(y, orc.faf_middle(x))), flush=True)
    for size in (56, 224):
        guarded(faf, size)

    def swin(name):
        f = mods[name]
        m = build(SwinTransformerBlock, f["ctor"])
        y = m(util.seeded_input(f["input_shape"], f["input_seed"]).cuda())
        print("%s: %.3e" % (name, util.maxabs(y, f["outputs"]["y"])), flush=True)
    for name in ("swin_s0", "swin_s3", "swin_t1_s3", "swin_res7"):
        guarded(swin, name)

    def sda(name):
        f = mods[name]
        m = build(SwinDAttention, f["ctor"])
        x1 = util.seeded_input(f["x1_shape"], f["input_seed"])
        x2 = util.seeded_input(f["x2_shape"], f["input_seed"] + 100)
        y, _ = m(x1.cuda(), x2.cuda())
        print("%s: %.3e" % (name, util.maxabs(y, f["outputs"]["y"])), flush=True)
    for name in ("sda_r3", "sda_r1"):
        guarded(sda, name)

    def cross(name):
        f = mods[name]
        m = build(CrossSwinBlock, f["ctor"])
        x1 = util.seeded_input(f["x1_shape"], f["input_seed"])
        x2 = x1 if f["ctor"]["last_view"] else util.seeded_input(f["x2_shape"], f["input_seed"] + 100)
        y, out = m(x1.cuda(), x2.cuda())
        print("%s: y %.3e out %.3e" % (name, util.maxabs(y, f["outputs"]["y"]), util.maxabs(out, f["outputs"]["out"])), flush=True)
    for name in ("cross_r3", "cross_r1", "cross_last"):
        guarded(cross, name)

    def merge():
        f = mods["merge"]
        m = build(PatchMerging, f["ctor"])
        print("merge: %.3e" % util.maxabs(m(util.seeded_input(f["input_shape"], f["input_seed"]).cuda()), f["outputs"]["y"]), flush=True)
    guarded(merge)

    def vit():
        f = mods["vit_block"]
        m = build(Block, f["ctor"])
        print("vit_block: %.3e" % util.maxabs(m(util.seeded_input(f["input_shape"], f["input_seed"]).cuda()), f["outputs"]["y"]), flush=True)
    guarded(vit)

    def tok():
        f = mods["tokenize"]
        m = CrossThreeViewTokenize(default_view_configs()).eval()
        util.load_seeded(m)
        out = m.cuda()(util.seeded_input(f["input_shape"], f["input_seed"]).cuda())
        print("tokenize:", ["%.3e" % util.maxabs(out[i].reshape(2, -1, out[i].shape[-1]), f["outputs"]["v%d" % i]) for i in range(3)], flush=True)
    guarded(tok)

    def dec():
        f = mods["decoder"]
        m = mumpy_b200.Decoder().eval()
        util.load_seeded(m)
        m = m.cuda()
        final_x, view_x, ff = util.decoder_inputs(1)
        lg, xf = m(final_x.cuda(), [[t.cuda() for t in st] for st in view_x], ff.cuda())
        print("decoder: logits %.3e x_feats %.3e" % (util.maxabs(lg, f["outputs"]["logits"]), util.maxabs(xf[:, :, ::4, ::4], f["outputs"]["x_feats_sub"])), flush=True)
    guarded(dec)


def e2e(model, mode, B):
    g = util.golden("e2e_b%d.pt" % B)
    x = util.seeded_input(g["input_shape"], g["input_seed"]).cuda()
    mumpy_b200.set_precision(mode)
    enc, dec = model
    with torch.no_grad():
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            n0 = ops.launch_count
            final_x, view_x, ff = enc(x)
            logits, feats = dec(final_x, view_x, ff)
            torch.cuda.synchronize()
            dt = time.time() - t0
    print("e2e %s B=%d: %.1f ms (eager, %d C-ABI launches)" % (mode, B, dt * 1e3, ops.launch_count - n0))
    print("  ffinfo %.3e" % util.maxabs(ff[:, :, ::4, ::4], g["ffinfo_sub"]))
    for s in range(4):
        print("  stage %d views:" % s, ["%.3e" % util.maxabs(view_x[s][v][:, :, ::16, :], g["view_sub"][s][v]) for v in range(3)])
    d = (logits.cpu() - g["logits"]).abs()
    print("  final_x %.3e  x_feats %.3e" % (util.maxabs(final_x, g["final_x"]), util.maxabs(feats[:, :, ::4, ::4], g["x_feats_sub"])))
    print("  logits max-abs %.3e mean-abs %.3e  mask identity %.5f" % (float(d.max()), float(d.mean()),
          float(((logits.cpu() > 0) == (g["logits"] > 0)).float().mean())), flush=True)


def main():
    print(torch.cuda.get_device_name(0), "abi", mumpy_b200._lib.load().mumpy_abi_version())
    section("GEMM fp32 (SIMT)")
    guarded(gemm_checks, False)
    mumpy_b200.set_precision("fp32")
    section("modules, fp32 mode vs reference captures")
    guarded(module_checks)
    section("end-to-end fp32")
    enc, dec = mumpy_b200.Encoder().eval(), mumpy_b200.Decoder().eval()
    util.load_seeded(enc)
    util.load_seeded(dec)
    model = (enc.cuda(), dec.cuda())
    for B in (1, 2):
        guarded(e2e, model, "fp32", B)
    if "--skip-bf16" in sys.argv:
        return
    section("GEMM bf16 (tcgen05)")
    ok = guarded(gemm_checks, True)
    if ok:
        section("end-to-end bf16")
        for B in (1, 2):
            guarded(e2e, model, "bf16", B)


if __name__ == "__main__":
    main()
