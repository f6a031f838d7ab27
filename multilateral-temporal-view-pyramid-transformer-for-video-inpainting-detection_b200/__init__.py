"""mumpy_b200 -- B200-native (sm_100a) implementation of Mumpy's inference forward.

The product package lives in the directory
`multilateral-temporal-view-pyramid-transformer-for-video-inpainting-detection_b200/` (not an importable
identifier); `import mumpy_b200` (repo root shim) resolves to it.  Layout:

  csrc/                 hand-written CUDA kernels + the C ABI (include/mumpy_b200.h) -> libmumpy_b200.so
  _lib.py, ops.py       ctypes binding and tensor-level wrappers (device memory / streams via PyTorch only)
  models/               mirrors of the reference's nn.Module classes (same names, signatures, state_dict keys)
  evaluate.py           clip-sharded evaluation: thresholded masks, per-clip F1/IoU counts, NCCL reduction
  build.py              nvcc recipe for the shared library
"""
from . import ops, streams  # noqa: F401
from .ops import check_f16_range, f16_overflowed, precision, set_precision  # noqa: F401
from .models.encoder.encoder import Encoder  # noqa: F401
from .models.decoder.decoder import Decoder  # noqa: F401



def forward(encoder, decoder, x):
    """test.py:94-95 as one call: `feats, views, dct = encoder(x); return decoder(feats, views, dct)` inside ONE stream region,
    so that the decoder's pyramid / frequency branches start as soon as their stage features exist and overlap the
    encoder's tail (stage 3 and the 12 small global blocks) instead of waiting for the encoder's join.  Same results, bit for
    bit, as the two separate calls (tests/test_gpu_e2e.py).  Returns (logits (B,1,S,S), x_feats (B,32,S,S))."""
    # The input conversion (non-contiguous or non-fp32 clips) must be enqueued BEFORE the region forks its lanes: the
    # frequency lane and the tokenizer lanes read x without any further hand-over.
    x = x.contiguous().float()
    with streams.region(x.device) as reg:
        reg.hold(x)
        final_x, view_x, ffinfo = encoder(x)
        return decoder(final_x, view_x, ffinfo)


__all__ = ["Encoder", "Decoder", "forward", "ops", "streams", "set_precision", "precision", "f16_overflowed", "check_f16_range"]
