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
from . import ops  # noqa: F401
from .ops import precision, set_precision  # noqa: F401
from .models.encoder.encoder import Encoder  # noqa: F401
from .models.decoder.decoder import Decoder  # noqa: F401

__all__ = ["Encoder", "Decoder", "ops", "set_precision", "precision"]
