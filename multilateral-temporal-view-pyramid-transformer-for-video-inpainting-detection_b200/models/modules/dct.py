"""Frequency branch (FAF) on libmumpy_b200 kernels.

Mirrors reference models/modules/dct.py: DCT_mat (:42-45), generate_filter (:48-49), Filter (:11-39), FAF (:56-79).
FAF is a full-frame 2-D DCT-II band split (not an 8x8 block DCT): Y_k = D^T (F_k o (D X D^T)) D for three bands
of i+j.  It has no parameters or buffers in the state_dict (the reference keeps plain tensor attributes).
"""
import math

import torch
import torch.nn as nn

from ... import ops
from ._packing import require_inference


def DCT_mat(size):
    return [[math.sqrt(1. / size) if i == 0 else math.sqrt(2. / size) * math.cos((j + 0.5) * math.pi * i / size)
             for j in range(size)] for i in range(size)]


def generate_filter(start, end, size):
    return [[0. if i + j > end or i + j < start else 1. for j in range(size)] for i in range(size)]


class Filter(nn.Module):
    """Band description holder (dct.py:11-39 with use_learnable=False, norm=False): keeps start <= i+j <= end."""

    def __init__(self, size, band_start, band_end, use_learnable=False, norm=False, fine_grain=False):
        super().__init__()
        if use_learnable or norm or fine_grain:
            raise NotImplementedError("only the fixed band filters FAF constructs are implemented")
        self.size = size
        self.band = (int(math.ceil(band_start)), int(math.floor(band_end)))

    def forward(self, x):
        raise RuntimeError("Filter is fused into the FAF kernel and is not called on its own")


class FAF(nn.Module):
    def __init__(self, size=224):
        super().__init__()
        self.fn = 3
        self.size = size
        self._dct = torch.tensor(DCT_mat(size), dtype=torch.float64).float()    # same rounding as dct.py:59
        low_filter = Filter(size, 0, size // 2.82)
        middle_filter = Filter(size, size // 2.82, size // 2)
        high_filter = Filter(size, size * 1, size * 2)
        self.filters = nn.ModuleList([low_filter, middle_filter, high_filter])

    def _dct_on(self, device):
        if self._dct.device != device:
            self._dct = self._dct.to(device)
        return self._dct

    def _split_operands(self, device):
        """[D_hi | D_lo | D_hi] and the same for D^T in the operand type of the current 16-bit mode (cached per device / mode)."""
        key = (str(device), ops.precision())
        cache = self.__dict__.setdefault("_split_cache", {})
        if key not in cache:
            d = self._dct_on(device)
            dt16 = ops.act_dtype()

            def cat(m):
                hi = m.to(dt16)
                lo = (m - hi.float()).to(dt16)
                return torch.cat([hi, lo, hi], dim=1).contiguous()
            cache[key] = (cat(d), cat(d.t().contiguous()))
            if not torch.cuda.is_current_stream_capturing():
                torch.cuda.current_stream(device).synchronize()       # cached across streams, like _packing.PackedModule
        return cache[key]

    def frame(self, x, frame=1):
        """x (B,T,3,S,S) -> FAF of one frame, (B,9,S,S), channel = band*3 + rgb."""
        require_inference(self)
        bands = [f.band for f in self.filters]
        if ops.tensor_cores() and self.size % 8 == 0:
            dcat, dtcat = self._split_operands(x.device)
            return ops.faf16(x.contiguous(), dcat, dtcat, bands, frame)
        return ops.faf(x.contiguous(), self._dct_on(x.device), bands, frame)

    def forward(self, x):
        """x (B,T,3,S,S) -> (B,T,9,S,S) like the reference; the encoder only needs frame 1 and calls frame()."""
        return torch.stack([self.frame(x, t) for t in range(x.shape[1])], dim=1)
