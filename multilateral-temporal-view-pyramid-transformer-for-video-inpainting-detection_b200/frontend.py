"""Device-side front end of the evaluation loop: the reference builds every 3-frame clip on
the host (universaldataloader.py:45-48: frames idx-1, idx, idx+1 clamped to the sequence), converts it with
ToTensor + Normalize (test.py:22-25) and uploads 1.8 MB of fp32 per clip.  Consecutive clips share two of their three
frames, so here every frame is uploaded ONCE as uint8 HWC (150 KB at 224x224) and the clips are assembled and normalised on
the GPU by mumpy_assemble_clips (bit-identical to the torchvision transform).  Frames may also be handed over at their native
resolution: ClipAssembler.from_native resizes them on the device with mumpy_resize_u8, bit-identical to the PIL resize of the
loader.
"""
import torch

from . import ops

MEAN = (0.4776, 0.479, 0.4465)          # test.py:23-24
STD = (0.230, 0.2085, 0.2324)


def clip_frame_indices(seq_lengths, length_clip=3):
    """(n_clips, length_clip) int32 frame indices into the concatenation of all sequences: one clip per frame, centred on it,
    neighbours clamped to the sequence (`max(0, min(n - 1, i)) for i in range(idx - k, idx + k + 1)`, k = length_clip // 2)."""
    k = length_clip // 2
    rows, base = [], 0
    for n in seq_lengths:
        idx = torch.arange(n).unsqueeze(1) + torch.arange(-k, k + 1).unsqueeze(0)
        rows.append(idx.clamp_(0, n - 1) + base)
        base += n
    return torch.cat(rows, 0).to(torch.int32) if rows else torch.zeros((0, length_clip), dtype=torch.int32)


class ClipAssembler:
    """Holds the uint8 frames of a split on the device and hands out normalised clip batches.

    frames: (n_frames, H, W, 3) uint8 (host or device; uploaded once), seq_lengths: frames per sequence in order."""

    def __init__(self, frames, seq_lengths, device, length_clip=3, mean=MEAN, std=STD):
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError("frames must be (n, H, W, 3) uint8")
        if int(sum(seq_lengths)) != frames.shape[0]:
            raise ValueError("seq_lengths do not add up to the number of frames")
        self.frames = frames.to(device, non_blocking=True).contiguous()
        self.index = clip_frame_indices(seq_lengths, length_clip).to(device)
        self.mean, self.std = mean, std

    @classmethod
    def from_native(cls, frames, seq_lengths, device, size=224, resample=ops.RESIZE_BICUBIC, chunk=64, **kw):
        """frames at their native resolution ((n, H0, W0, 3) uint8, e.g. 480 x 854 DAVIS frames): uploaded as they are and resized
        to size x size on the device exactly as the loader's `img.resize(self.inputRes)` does on the host
        (universaldataset.py:68-79; `resample` = PIL filter code: bicubic is Pillow >= 7's default, nearest pillow 4.0.0's)."""
        small = [ops.resize_u8(frames[i:i + chunk].to(device, non_blocking=True), size, size, resample) for i in range(0, frames.shape[0], chunk)]
        return cls(torch.cat(small, 0), seq_lengths, device, **kw)

    def __len__(self):
        return self.index.shape[0]

    def batch(self, lo, hi, out=None):
        """Clips [lo, hi) as (hi-lo, T, 3, H, W) fp32, normalised."""
        return ops.assemble_clips(self.frames, self.index[lo:hi].contiguous(), self.mean, self.std, out=out)
