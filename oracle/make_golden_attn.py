"""TEST INFRASTRUCTURE ONLY -- attention maps of the UNMODIFIED reference (/root/reference through oracle/ref_harness.py, CPU, fp32):
the second result of SwinDAttention.forward (deformableAttention.py:399,405), CVAModule.forward(return_attention=True)
(multiTemporalViewEncoder.py:134-137) and Block.forward(return_attention=True) (blocks.py:86-89), with the same key-seeded weights and
seeded inputs as the module fixtures of make_golden.py.  Run in the build container:

    python oracle/make_golden_attn.py        # writes tests/golden/attn_maps.pt and ATTN_PIN_REPORT.json

and pins the oracle's `return_attn` / `return_attention` outputs against them.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mumpy_oracle as orc          # noqa: E402
from oracle import ref_harness as rh            # noqa: E402
from oracle.make_golden import GOLD, load_seeded, maxabs, seeded_input      # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    ref = rh.load()
    out, report = {}, {}
    with torch.no_grad():
        for name, dim, heads, n1, ratio, seed in (("sda_r3", 96, 3, 4, 3, 31), ("sda_r1", 192, 6, 3, 1, 32)):
            m = ref.datt.SwinDAttention(dim, heads, 0.0, n_groups=3).eval()
            sd = load_seeded(m)
            x1 = seeded_input((n1, 49, dim), seed)
            x2 = seeded_input((n1 * ratio, 49, dim), seed + 100)
            y, attn = m(x1, x2)
            yo, attno = orc.swin_dattention(sd, "", x1, x2, heads, return_attn=True)
            report[name] = {"y": maxabs(y, yo), "attn": maxabs(attn, attno), "attn_shape": list(attn.shape)}
            out[name] = {"input_seed": seed, "x1_shape": list(x1.shape), "x2_shape": list(x2.shape),
                         "ctor": dict(dim1=dim, n_heads=heads, attn_drop=0.0, n_groups=3), "attn": attn.clone()}
        # CVAModule(return_attention=True) returns the same map
        m = ref.mtv.CVAModule(96, 3).eval()
        sd = load_seeded(m)
        x1, x2 = seeded_input((4, 49, 96), 33), seeded_input((12, 49, 96), 133)
        attn = m(x1, x2, return_attention=True)
        _, attno = orc.swin_dattention(sd, "crossattn.", x1, x2, 3, return_attn=True)
        report["cva_module"] = {"attn": maxabs(attn, attno), "attn_shape": list(attn.shape)}
        out["cva_module"] = {"input_seed": 33, "x1_shape": list(x1.shape), "x2_shape": list(x2.shape), "ctor": dict(dim1=96, num_heads=3),
                             "attn": attn.clone()}
        m = ref.blocks.Block(128, 4, 256, 0.0, 0.0).eval()
        sd = load_seeded(m)
        x = seeded_input((10, 3, 128), 61)
        attn = m(x, return_attention=True)
        attno = orc.vit_block(sd, "", x, 4, return_attention=True)
        report["vit_block"] = {"attn": maxabs(attn, attno), "attn_shape": list(attn.shape)}
        out["vit_block"] = {"input_seed": 61, "input_shape": list(x.shape), "ctor": dict(dim=128, heads=4, mlp_dim=256, dropout=0.0, drop_path=0.0),
                            "attn": attn.clone()}
    torch.save(out, os.path.join(GOLD, "attn_maps.pt"))
    json.dump(report, open(os.path.join(GOLD, "ATTN_PIN_REPORT.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
