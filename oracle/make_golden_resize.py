"""TEST INFRASTRUCTURE ONLY -- pins oracle/pil_resample.py against Pillow itself (the third-party library whose
`Image.resize` the reference's loader calls, dataloaders/universaldataset.py:68-79) and writes the fixtures the CPU and GPU
tests compare with:

    python oracle/make_golden_resize.py      # writes tests/golden/resize_pil.npz, prints the pin report

Inputs are seeded (numpy default_rng(seed) uint8 noise, blurred along both axes so that neighbouring pixels correlate like an
image, plus a saturated block so that the clip-to-[0,255] path of the cubic filter's overshoot is exercised); only the seeds,
shapes and PIL's outputs are stored.
"""
import json
import os
import sys

import numpy as np
import PIL
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pil_resample as pr      # noqa: E402

CASES, seeded_image = pr.CASES, pr.seeded_image


def pil_resize(img, out_h, out_w, flt):
    im = Image.fromarray(img[..., 0] if img.shape[2] == 1 else img)
    out = np.asarray(im.resize((out_w, out_h), flt))
    return out[..., None] if out.ndim == 2 else out


def main():
    store, report = {}, {"pillow": PIL.__version__, "cases": []}
    for seed, h, w, c, oh, ow in CASES:
        img = seeded_image(seed, h, w, c)
        for name, flt, fn in (("bicubic", Image.BICUBIC, pr.resize_bicubic_u8), ("nearest", Image.NEAREST, pr.resize_nearest_u8)):
            ref = pil_resize(img, oh, ow, flt)
            mine = fn(img, oh, ow)
            n_bad = int((ref != mine).sum())
            report["cases"].append({"case": [seed, h, w, c, oh, ow], "filter": name, "mismatching_bytes": n_bad})
            assert n_bad == 0, (seed, name, n_bad)
            if h * w <= 120 * 214 or name == "bicubic" and seed == 3:       # keep the fixture file small; the rest is re-derivable
                store["%s_%d" % (name, seed)] = ref
        # the default filter of the installed Pillow is the bicubic one
        assert np.array_equal(pil_resize(img, oh, ow, None), pil_resize(img, oh, ow, Image.BICUBIC))
    store["cases"] = np.array(CASES, dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_pil.npz"), **store)
    with open(os.path.join(ROOT, "tests", "golden", "RESIZE_PIN_REPORT.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
