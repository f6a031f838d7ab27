"""Development aid: where does a patched-resolution forward first deviate from the oracle (fp32 mode)?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mumpy_b200  # noqa: E402
from mumpy_b200 import ops  # noqa: E402
from oracle import mumpy_oracle as orc  # noqa: E402
from tests import util  # noqa: E402

size, ws = int(sys.argv[1]), int(sys.argv[2])
res = tuple(size // d for d in (4, 8, 16, 32))
mumpy_b200.set_precision("fp32")
enc = mumpy_b200.Encoder(img_size=size, window_size=ws).eval()
sd = util.load_seeded(enc.base)
osd = {"base." + k: v for k, v in sd.items()}
enc = enc.cuda()
x = util.seeded_input((1, 3, 3, size, size), 5)
cfg = orc.default_config(res=res, ws=ws, img=size)
with torch.no_grad():
    toks = enc.base.tokenize(x.cuda())
    otoks = orc.tokenize(osd, x)
    for v in range(3):
        print("tokenize v%d" % v, util.maxabs(toks[v].reshape(otoks[v].shape), otoks[v]))
    # stage 0, block 0 of view 3 alone (last view: plain attention + mlp) and plain block 1 (shifted)
    blk0, blk1 = enc.base.layers.layers[0].blocks[0], enc.base.layers.layers[0].blocks[1]
    x3 = otoks[2].reshape(1, -1, 128)
    h3, _, _ = blk0.block3.attn_phase(x3.cuda(), need_out=False, need_out_fp32=True)
    y3 = blk0.block3.tail_phase(h3, None)
    o3, _ = orc.cross_swin_block(osd, "base.layers.layers.0.blocks.0.block3.", x3, x3, 3 * res[0], 3 * res[0], res[0], 4, ws, True)
    print("s0 b0 view3", util.maxabs(y3, o3))
    z3 = blk1.block3(o3.cuda())
    oz3 = orc.swin_block(osd, "base.layers.layers.0.blocks.1.block3.", o3, 3 * res[0], res[0], 4, ws, ws // 2)
    print("s0 b1 view3 (shifted)", util.maxabs(z3, oz3))
    final_x, view_x, ff = enc(x.cuda())
    o_final, o_view, o_ff = orc.encoder_forward(osd, x, cfg)
    for s in range(4):
        for v in range(3):
            print("stage %d view %d" % (s, v), util.maxabs(view_x[s][v], o_view[s][v]))
    print("final", util.maxabs(final_x, o_final))
