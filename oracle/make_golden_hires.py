"""TEST INFRASTRUCTURE ONLY -- fixtures for the patched-resolution configurations (BASELINE.json configs[3], SURVEY 8(d)
"Config 4", appendix A10): the reference cannot run 512x512 as shipped (window 7 does not divide 128 tokens, FAF() is fixed
to 224, SwinDAttention.ws = 7, the final rearrange assumes 7x7), so -- exactly as the survey verified -- the UNMODIFIED
reference classes are instantiated with window_size 8, input_resolution size/4 .. size/32, FAF(size), SwinDAttention.ws = 8,
Decoder(shape=...) and the final rearrange with h = w = size/32.

    python oracle/make_golden_hires.py      # writes tests/golden/e2e_256w8_b1.pt, e2e_512w8_b1.pt and extends PIN_REPORT.json
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
from oracle.make_golden import GOLD, load_seeded, maxabs, seeded_input, sub_tokens      # noqa: E402


def build_reference(size, ws):
    ref = rh.load()
    res = tuple(size // d for d in (4, 8, 16, 32))
    cfgs = rh.view_configs(ref, window_size=ws, res=res)
    genc = rh._ConfigDict({'num_heads': 12, 'mlp_dim': 3072, 'num_layers': 12, 'hidden_size': 768,
                           'merge_axis': 'channel', 'num_frames': 3})
    enc = ref.mtv.ThreeViewSwinTransformer(view_configs=cfgs, input_token_temporal_dims=[1, 1, 3], global_encoder_config=genc,
                                           depths=[2, 2, 18, 2]).eval()
    enc.faf = ref.dct.FAF(size)
    for m in enc.modules():
        if isinstance(m, ref.datt.SwinDAttention):
            m.ws = ws
    dec = ref.dec.Decoder(shape=list(res)).eval()
    return enc, dec, res


def main():
    torch.set_num_threads(os.cpu_count())
    report_path = os.path.join(GOLD, "PIN_REPORT.json")
    report = json.load(open(report_path)) if os.path.exists(report_path) else {}
    for size, ws, seed in ((256, 8, 2256), (512, 8, 2512)):
        enc, dec, res = build_reference(size, ws)
        enc_sd = load_seeded(enc)              # keys without the Encoder wrapper's "base." prefix
        dec_sd = load_seeded(dec)
        x = seeded_input((1, 3, 3, size, size), seed)
        with torch.no_grad():
            t0 = time.time()
            final, view_x, ffinfo = enc(x)
            side = res[3]
            final_x = final.reshape(1, side, side, -1).permute(0, 3, 1, 2).contiguous()     # encoder.py:16-17 with h = w = size/32
            logits, feats = dec(final_x, view_x, ffinfo)
            t_ref = time.time() - t0
            cfg = orc.default_config(res=res, ws=ws, img=size)
            osd = {"base." + k: v for k, v in enc_sd.items()}
            t0 = time.time()
            o_final, o_view, o_ff = orc.encoder_forward(osd, x, cfg)
            o_logits, o_feats = orc.decoder_forward(dec_sd, o_final, o_view, o_ff, res)
            t_orc = time.time() - t0
        rep = {"ref_seconds": t_ref, "oracle_seconds": t_orc, "logits": maxabs(logits, o_logits), "x_feats": maxabs(feats, o_feats),
               "final_x": maxabs(final_x, o_final), "ffinfo": maxabs(ffinfo, o_ff),
               "mask_flips": int(((logits > 0) != (o_logits > 0)).sum()),
               "logit_stats": [float(logits.mean()), float(logits.std()), float((logits > 0).float().mean())]}
        report["e2e_%dw%d_b1" % (size, ws)] = rep
        print(size, ws, rep)
        torch.save({"size": size, "window": ws, "res": list(res), "input_seed": seed, "input_shape": list(x.shape),
                    "logits": logits.contiguous(), "final_x": final_x, "x_feats_sub": feats[:, :, ::8, ::8].contiguous(),
                    "ffinfo_sub": ffinfo[:, :, ::8, ::8].contiguous(),
                    "view_sub": [[sub_tokens(v, 64) for v in st] for st in view_x]},
                   os.path.join(GOLD, "e2e_%dw%d_b1.pt" % (size, ws)))
    with open(report_path, "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
