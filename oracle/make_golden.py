"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/* by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) on CPU in fp32 with key-seeded weights
(oracle/weights.py) and seeded inputs.  Run in the build container:

    python oracle/make_golden.py            # writes tests/golden/*.pt, manifest.json, PIN_REPORT.json

It also pins the oracle restatement (oracle/mumpy_oracle.py) against every fixture and records the
max-abs differences in tests/golden/PIN_REPORT.json.  Large tensors are stored as strided subsamples
(the slicing recipe is stored beside them) to keep the fixtures small.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mumpy_oracle as orc          # noqa: E402
from oracle import ref_harness as rh            # noqa: E402
from oracle import weights as wts               # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def seeded_input(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def sub_tokens(t, stride=16):
    """(B,1,L,C) -> tokens ::stride"""
    return t[:, :, ::stride, :].contiguous()


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


def load_seeded(module, seed=0):
    sd = wts.fill_state_dict(module.state_dict(), seed)
    module.load_state_dict(sd, strict=True)
    return sd


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    ref = rh.load()
    report = {}

    # ------------------------------------------------------------------ end-to-end
    enc = rh.ReferenceEncoder().eval()
    dec = rh.build_reference_decoder()
    enc_sd = load_seeded(enc)
    dec_sd = load_seeded(dec)
    manifest = {
        "encoder": {k: list(v.shape) for k, v in enc_sd.items() if not wts.is_buffer_key(k)},
        "decoder": {k: list(v.shape) for k, v in dec_sd.items() if not wts.is_buffer_key(k)},
        "encoder_buffers": {k: list(v.shape) for k, v in enc_sd.items() if wts.is_buffer_key(k)},
    }
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)

    for B, seed in ((1, 1234), (2, 1235)):
        x = seeded_input((B, 3, 3, 224, 224), seed)
        with torch.no_grad():
            t0 = time.time()
            final_x, view_x, ffinfo = enc(x)
            logits, feats = dec(final_x, view_x, ffinfo)
            t_ref = time.time() - t0
            t0 = time.time()
            o_final, o_view, o_ff = orc.encoder_forward(enc_sd, x)
            o_logits, o_feats = orc.decoder_forward(dec_sd, o_final, o_view, o_ff)
            t_orc = time.time() - t0
        rep = {
            "ref_seconds": t_ref, "oracle_seconds": t_orc,
            "logits": maxabs(logits, o_logits), "x_feats": maxabs(feats, o_feats),
            "final_x": maxabs(final_x, o_final), "ffinfo": maxabs(ffinfo, o_ff),
            "mask_flips": int(((logits > 0) != (o_logits > 0)).sum()),
            "logits_mean": float(logits.mean()), "logits_std": float(logits.std()),
            "frac_pos": float((logits > 0).float().mean()),
        }
        for s in range(4):
            for v in range(3):
                rep["view_s%d_v%d" % (s, v)] = maxabs(view_x[s][v], o_view[s][v])
        report["e2e_b%d" % B] = rep
        print("e2e B=%d" % B, json.dumps(rep))
        torch.save({
            "input_seed": seed, "input_shape": [B, 3, 3, 224, 224], "weight_seed": 0,
            "logits": logits.clone(), "final_x": final_x.clone(),
            "x_feats_sub": feats[:, :, ::4, ::4].clone(), "x_feats_slice": "[:, :, ::4, ::4]",
            "ffinfo_sub": ffinfo[:, :, ::4, ::4].clone(), "ffinfo_slice": "[:, :, ::4, ::4]",
            "view_sub": [[sub_tokens(t) for t in st] for st in view_x], "view_slice": "[:, :, ::16, :]",
        }, os.path.join(GOLD, "e2e_b%d.pt" % B))
        if B == 1:
            # isolated decoder fixture from the same activations: lets decoder parity be checked alone
            pass

    # ------------------------------------------------------------------ module-level fixtures
    mods = {}

    def record(name, out_ref, out_orc, payload):
        d = {k: maxabs(a, b) for k, (a, b) in zip(out_ref.keys(), zip(out_ref.values(), out_orc.values()))}
        report[name] = d
        print(name, json.dumps(d))
        payload["outputs"] = {k: v.clone() for k, v in out_ref.items()}
        mods[name] = payload

    with torch.no_grad():
        # FAF at a small size and at 224 (dct.py:56-79)
        for size, seed in ((56, 11), (224, 12)):
            faf = ref.dct.FAF(size)
            x = seeded_input((2, 3, 3, size, size), seed)
            y = faf(x)[:, 1]
            record("faf_%d" % size, {"y": y}, {"y": orc.faf_middle(x)},
                   {"input_seed": seed, "input_shape": list(x.shape), "size": size})
            if size == 224:
                mods["faf_224"]["outputs"] = {"y_sub": y[:, :, ::4, ::4].clone()}
                mods["faf_224"]["slice"] = "[:, :, ::4, ::4]"

        # SwinTransformerBlock (swinTransformer.py:185-307), stacked frames, shifted and not
        for name, dim, H, heads, T, shift, seed in (("swin_s0", 64, 14, 2, 3, 0, 21), ("swin_s3", 64, 14, 2, 3, 3, 22),
                                                    ("swin_t1_s3", 96, 14, 3, 1, 3, 23), ("swin_res7", 64, 7, 2, 3, 3, 24)):
            m = ref.swin.SwinTransformerBlock(dim, (H, H), heads, window_size=7, shift_size=shift, temporal_dim=T).eval()
            sd = load_seeded(m)
            x = seeded_input((2, T * H * H, dim), seed)
            y = m(x)
            eff_shift = 0 if H <= 7 else shift
            yo = orc.swin_block(sd, "", x, T * H, H, heads, min(7, H), eff_shift)
            record(name, {"y": y}, {"y": yo}, {"input_seed": seed, "input_shape": list(x.shape),
                                               "ctor": dict(dim=dim, input_resolution=[H, H], num_heads=heads, window_size=7,
                                                            shift_size=shift, temporal_dim=T)})

        # SwinDAttention (deformableAttention.py:218-405): ratio 3 and ratio 1
        for name, dim, heads, n1, ratio, seed in (("sda_r3", 96, 3, 4, 3, 31), ("sda_r1", 192, 6, 3, 1, 32)):
            m = ref.datt.SwinDAttention(dim, heads, 0.0, n_groups=3).eval()
            sd = load_seeded(m)
            x1 = seeded_input((n1, 49, dim), seed)
            x2 = seeded_input((n1 * ratio, 49, dim), seed + 100)
            y, _ = m(x1, x2)
            yo = orc.swin_dattention(sd, "", x1, x2, heads)
            record(name, {"y": y}, {"y": yo}, {"input_seed": seed, "x1_shape": list(x1.shape), "x2_shape": list(x2.shape),
                                               "ctor": dict(dim1=dim, n_heads=heads, attn_drop=0.0, n_groups=3)})

        # CrossSwinBlock (multiTemporalViewEncoder.py:142-291): v2<-v3 style (ratio 3) and last_view
        for name, d1, d2, H, heads, T1, T2, last, seed in (("cross_r3", 96, 128, 14, 3, 1, 3, False, 41),
                                                           ("cross_r1", 96, 96, 14, 3, 1, 1, False, 42),
                                                           ("cross_last", 128, 128, 14, 4, 3, 3, True, 43)):
            m = ref.mtv.CrossSwinBlock(d1, d2, (H, H), heads, window_size=7, last_view=last, temporal_dims=T1).eval()
            sd = load_seeded(m)
            x1 = seeded_input((2, T1 * H * H, d1), seed)
            x2 = x1 if last else seeded_input((2, T2 * H * H, d2), seed + 100)
            y, out = m(x1, x2)
            yo, outo = orc.cross_swin_block(sd, "", x1, x2, T1 * H, T2 * H, H, heads, 7, last)
            record(name, {"y": y, "out": out}, {"y": yo, "out": outo},
                   {"input_seed": seed, "x1_shape": list(x1.shape), "x2_shape": list(x2.shape),
                    "ctor": dict(dim1=d1, dim2=d2, input_resolution=[H, H], num_heads=heads, window_size=7,
                                 last_view=last, temporal_dims=T1)})

        # PatchMerging (swinTransformer.py:328-367)
        m = ref.swin.PatchMerging((3 * 14, 14), 64).eval()
        sd = load_seeded(m)
        x = seeded_input((2, 3 * 14 * 14, 64), 51)
        record("merge", {"y": m(x)}, {"y": orc.patch_merging(sd, "", x, 42, 14)},
               {"input_seed": 51, "input_shape": list(x.shape), "ctor": dict(input_resolution=[42, 14], dim=64)})

        # ViT Block (blocks.py:77-92) on 3 temporal tokens
        m = ref.blocks.Block(128, 4, 256, 0.0, 0.0).eval()
        sd = load_seeded(m)
        x = seeded_input((10, 3, 128), 61)
        record("vit_block", {"y": m(x)}, {"y": orc.vit_block(sd, "", x, 4)},
               {"input_seed": 61, "input_shape": list(x.shape), "ctor": dict(dim=128, heads=4, mlp_dim=256, dropout=0.0, drop_path=0.0)})

        # Tokenizer (multiTemporalViewEncoder.py:574-618)
        cfgs = rh.view_configs(ref)
        m = ref.mtv.CrossThreeViewTokenize(cfgs).eval()
        sd = load_seeded(m)
        x = seeded_input((2, 3, 3, 56, 56), 71)
        yr = m(x)
        yo = orc.tokenize(sd, x, pre="")
        record("tokenize", {"v%d" % i: yr[i].reshape(2, -1, yr[i].shape[-1]) for i in range(3)},
               {"v%d" % i: yo[i] for i in range(3)}, {"input_seed": 71, "input_shape": list(x.shape)})

        # Decoder alone on seeded activations (decoder.py:183-225)
        B = 1
        final_x = seeded_input((B, 2304, 7, 7), 81)
        ff = seeded_input((B, 9, 224, 224), 82)
        view_x = []
        for s in range(4):
            h = (56, 28, 14, 7)[s]
            view_x.append([seeded_input((B, 1, orc.VIEW_T[v] * h * h, orc.VIEW_DIMS[v][s]), 83 + 3 * s + v) for v in range(3)])
        lg, xf = dec(final_x, view_x, ff)
        lgo, xfo = orc.decoder_forward(dec_sd, final_x, view_x, ff)
        record("decoder", {"logits": lg, "x_feats": xf}, {"logits": lgo, "x_feats": xfo},
               {"seeds": "final_x 81, ffinfo 82, view_x[s][v] 83+3s+v", "weight_seed": 0})
        mods["decoder"]["outputs"] = {"logits": lg.clone(), "x_feats_sub": xf[:, :, ::4, ::4].clone()}

    torch.save(mods, os.path.join(GOLD, "modules.pt"))
    with open(os.path.join(GOLD, "PIN_REPORT.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    worst = max(v for d in report.values() for k, v in d.items()
                if isinstance(v, float) and not k.endswith("seconds") and not k.startswith(("logits_", "frac_")))
    print("worst oracle-vs-reference max-abs:", worst)


if __name__ == "__main__":
    main()
