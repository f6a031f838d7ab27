"""Clip assembly front end: index logic and ToTensor + Normalize arithmetic.
CPU: the oracle against torchvision's own transform objects (what test.py:22-25 composes) and the frame-index table
against the reference's list comprehension (universaldataloader.py:45-48).  GPU: mumpy_assemble_clips bit-exact vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import mumpy_oracle as orc
from tests import util


def _frames(n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, H, W, 3), generator=g, dtype=torch.uint8)


@pytest.mark.parametrize("lengths", [[1], [2], [1, 2, 5, 7], [3, 3, 3]])
def test_clip_frame_indices_match_reference_comprehension(lengths):
    from mumpy_b200 import frontend
    idx = frontend.clip_frame_indices(lengths, 3)
    want, base = [], 0
    for n in lengths:
        for i in range(n):
            want.append([base + max(0, min(n - 1, j)) for j in range(i - 1, i + 2)])     # universaldataloader.py:47
        base += n
    assert idx.dtype == torch.int32 and idx.tolist() == want


def test_oracle_transform_matches_torchvision():
    tv = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    frames = _frames(4, 24, 20, 3)
    tf = tv.Compose([tv.ToTensor(), tv.Normalize(mean=[0.4776, 0.479, 0.4465], std=[0.230, 0.2085, 0.2324])])
    clips = orc.assemble_clips(frames, [4])
    for i in range(4):
        ref = tf(Image.fromarray(frames[i].numpy()))
        assert torch.equal(clips[i, 1], ref)                      # centre frame of clip i is frame i
    assert torch.equal(clips[0, 0], clips[0, 1]) and torch.equal(clips[3, 2], clips[3, 1])      # edge clamping


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,lengths", [(224, 224, [5, 3]), (24, 20, [1, 2, 4])])
def test_assemble_clips_bit_exact(H, W, lengths):
    from mumpy_b200 import frontend
    frames = _frames(sum(lengths), H, W, 11)
    ref = orc.assemble_clips(frames, lengths)
    asm = frontend.ClipAssembler(frames, lengths, torch.device("cuda", 0))
    assert len(asm) == sum(lengths)
    out = asm.batch(0, len(asm))
    assert torch.equal(out.cpu(), ref)
    part = asm.batch(2, 4)
    assert torch.equal(part.cpu(), ref[2:4])


# ---------------------------------------------------------------- frame resize: PIL img.resize on the device
GOLD_RESIZE = np.load(__file__.rsplit("/", 1)[0] + "/golden/resize_pil.npz")


def test_resize_oracle_matches_pil_fixtures():
    """oracle/pil_resample.py against outputs of Pillow itself (tests/golden/resize_pil.npz, oracle/make_golden_resize.py)."""
    from oracle import pil_resample as pr
    checked = 0
    for seed, h, w, c, oh, ow in pr.CASES:
        img = pr.seeded_image(seed, h, w, c)
        for name, fn in (("bicubic", pr.resize_bicubic_u8), ("nearest", pr.resize_nearest_u8)):
            key = "%s_%d" % (name, seed)
            if key in GOLD_RESIZE and h * w <= 120 * 214:               # (the full-size case is covered on the GPU)
                assert np.array_equal(fn(img, oh, ow), GOLD_RESIZE[key]), key
                checked += 1
    assert checked >= 12


def test_resize_oracle_matches_installed_pillow():
    """Same check against the Pillow of this environment, when there is one (bicubic is its default filter)."""
    Image = pytest.importorskip("PIL.Image")
    from oracle import pil_resample as pr
    img = pr.seeded_image(21, 60, 107, 3)
    assert np.array_equal(pr.resize_bicubic_u8(img, 28, 28), np.asarray(Image.fromarray(img).resize((28, 28))))
    assert np.array_equal(pr.resize_nearest_u8(img, 28, 28), np.asarray(Image.fromarray(img).resize((28, 28), Image.NEAREST)))


@pytest.mark.parametrize("a,b", [(854, 224), (480, 224), (20, 45), (224, 224), (7, 3), (5, 2), (1920, 224), (224, 512)])
def test_resize_taps_of_the_library_match_the_oracle(a, b):
    """mumpy_resize_taps is host arithmetic (double precision, Pillow's operation order): bit-identical tables, no GPU needed."""
    from mumpy_b200 import ops
    from oracle import pil_resample as pr
    bounds, coefs, k = ops.resize_taps(a, b, ops.RESIZE_BICUBIC)
    rb, rc, rk = pr.bicubic_coeffs(a, b)
    assert k == rk and np.array_equal(bounds.numpy(), rb) and np.array_equal(coefs.numpy(), rc)
    idx, _, _ = ops.resize_taps(a, b, ops.RESIZE_NEAREST)
    assert np.array_equal(idx.numpy(), pr.nearest_index(a, b))


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(10))
def test_resize_u8_bit_exact(case):
    """mumpy_resize_u8 vs the oracle (and vs Pillow's stored output where the fixture holds it), both filters, batch of 2."""
    from mumpy_b200 import ops
    from oracle import pil_resample as pr
    seed, h, w, c, oh, ow = pr.CASES[case]
    imgs = np.stack([pr.seeded_image(seed, h, w, c), pr.seeded_image(seed + 100, h, w, c)])
    dev = torch.from_numpy(imgs).cuda()
    for name, code, fn in (("bicubic", ops.RESIZE_BICUBIC, pr.resize_bicubic_u8), ("nearest", ops.RESIZE_NEAREST, pr.resize_nearest_u8)):
        out = ops.resize_u8(dev, oh, ow, code).cpu().numpy()
        assert out.shape == (2, oh, ow, c)
        for i in range(2):
            assert np.array_equal(out[i], fn(imgs[i], oh, ow)), (name, i)
        key = "%s_%d" % (name, seed)
        if key in GOLD_RESIZE:
            assert np.array_equal(out[0], GOLD_RESIZE[key]), key


@pytest.mark.gpu
def test_clip_assembler_from_native_resolution():
    """Frames handed over at 480p: resized on the device, then identical to assembling the host-resized frames."""
    from mumpy_b200 import frontend
    from oracle import pil_resample as pr
    lengths = [3, 2]
    native = np.stack([pr.seeded_image(40 + i, 120, 214, 3) for i in range(5)])
    small = torch.from_numpy(np.stack([pr.resize_bicubic_u8(f, 56, 56) for f in native]))
    ref = orc.assemble_clips(small, lengths)
    asm = frontend.ClipAssembler.from_native(torch.from_numpy(native), lengths, torch.device("cuda", 0), size=56, chunk=2)
    assert torch.equal(asm.frames.cpu(), small)
    assert torch.equal(asm.batch(0, 5).cpu(), ref)
