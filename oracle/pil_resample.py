"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu legs, never by the product).

CPU restatement of the frame resize the reference's loader applies before ToTensor/Normalize
(dataloaders/universaldataset.py:68-79: `img.resize(self.inputRes)` on PIL images, inputRes = (224, 224), test.py:32), i.e. of
`PIL.Image.resize(size)` with its default filter.  The arithmetic lives in a third-party dependency that is NOT under
/root/reference: Pillow.  requirements.txt:9 pins pillow==4.0.0, whose default filter is NEAREST; Pillow >= 7 (and the 12.2.0
installed in this image) defaults to BICUBIC -- both are restated here:

  * BICUBIC, 8 bits per channel: Pillow's two-pass separable convolution resampler (src/libImaging/Resample.c:
    bicubic_filter (a = -0.5, support 2), precompute_coeffs, normalize_coeffs_8bpc (22 fractional bits),
    ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc, ImagingResampleInner): the horizontal pass runs first and its
    result is rounded to uint8 before the vertical pass; when down-scaling the filter support grows with the scale factor.
  * NEAREST: the affine scaler (src/libImaging/Geometry.c: ImagingScaleAffine): source index = (int)(start + scale / 2 +
    k * scale) with the position accumulated by repeated addition in double precision.

Pinned bit-exactly against PIL 12.2.0 in this container: tests/golden/resize_pil.npz (made by oracle/make_golden_resize.py).
All arithmetic is float64 / int64 numpy in the same operation order as the C source (no fused multiply-add).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are fixed point with 22 fractional bits


def _bicubic_filter(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = (0, in_size)).
    Returns (bounds int32 (out_size, 2) = [first source index, tap count], coefs int32 (out_size, ksize), ksize)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coefs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            coefs[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coefs, ksize


def _pass(img, bounds, coefs, axis):
    """One 8bpc pass along `axis` (0 = vertical, 1 = horizontal) of an (H, W, C) uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for k in range(n):
            acc += src[lo + k] * int(coefs[o, k])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, out_h, out_w):
    """(H, W, C) or (H, W) uint8 -> (out_h, out_w[, C]) uint8, equal to np.asarray(Image.fromarray(img).resize((out_w, out_h),
    Image.BICUBIC)) (ImagingResampleInner: horizontal pass first, only if the width changes; then vertical)."""
    x = img[..., None] if img.ndim == 2 else img
    if x.shape[1] != out_w:
        b, c, _ = bicubic_coeffs(x.shape[1], out_w)
        x = _pass(x, b, c, 1)
    if x.shape[0] != out_h:
        b, c, _ = bicubic_coeffs(x.shape[0], out_h)
        x = _pass(x, b, c, 0)
    x = np.ascontiguousarray(x)
    return x[..., 0] if img.ndim == 2 else x


def nearest_index(in_size, out_size):
    """ImagingScaleAffine's source index table for one axis: position accumulated by repeated addition."""
    scale = float(in_size) / out_size
    pos = 0.0 + scale * 0.5
    idx = np.zeros(out_size, dtype=np.int32)
    for o in range(out_size):
        i = -1 if pos < 0.0 else int(pos)
        idx[o] = min(max(i, 0), in_size - 1)         # (always inside for a whole-image box)
        pos += scale
    return idx


def resize_nearest_u8(img, out_h, out_w):
    """Equal to np.asarray(Image.fromarray(img).resize((out_w, out_h), Image.NEAREST))."""
    if img.shape[0] == out_h and img.shape[1] == out_w:
        return img.copy()
    return np.ascontiguousarray(img[nearest_index(img.shape[0], out_h)][:, nearest_index(img.shape[1], out_w)])


# ---- seeded test images shared by the fixture script and the tests (only seeds, shapes and PIL's outputs are stored)
# (seed, in_h, in_w, channels, out_h, out_w): DAVIS 480p aspect at reduced and full size, up-scaling, one-axis-only, identity
CASES = [(1, 48, 85, 3, 22, 22), (2, 120, 214, 3, 56, 56), (3, 480, 854, 3, 224, 224), (4, 20, 31, 3, 45, 64), (5, 224, 300, 3, 224, 224),
         (6, 300, 224, 3, 224, 224), (7, 64, 64, 3, 64, 64), (8, 97, 131, 1, 224, 224), (9, 480, 854, 1, 224, 224), (10, 7, 5, 3, 3, 2)]


def seeded_image(seed, h, w, c):
    """uint8 noise blurred along both axes (neighbouring pixels correlate like an image) with a saturated and a black block, so
    that the clip-to-[0,255] path of the cubic filter's overshoot is exercised."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, size=(h, w, c)).astype(np.float64)
    for ax in (0, 1):
        x = (x + np.roll(x, 1, ax) + np.roll(x, 2, ax) + np.roll(x, 3, ax)) / 4.0
    x = (x - x.min()) / max(x.max() - x.min(), 1e-9) * 255.0
    x[h // 4:h // 2, w // 4:w // 2] = 255.0
    x[h // 2:h // 2 + max(h // 8, 1), w // 2:w // 2 + max(w // 8, 1)] = 0.0
    return np.round(x).astype(np.uint8)
