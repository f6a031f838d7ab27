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
