"""TEST INFRASTRUCTURE ONLY -- deterministic, key-seeded "random-init" weights.

The reference's own constructor init depends on the global RNG stream and construction order, and
zero-initialises the deformable branch's proj_out (deformableAttention.py:308-309), which would leave
a9 numerically dead (SURVEY finding 5).  For parity every tensor is instead drawn from a generator
seeded by crc32(key), so the reference (make_golden.py), the oracle and the CUDA modules can all be
given bit-identical state_dicts from nothing but the key/shape manifest (tests/golden/manifest.json).

Scales are "trained-like": weights ~ N(0, gain^2/fan_in), norm scales ~ 1 + 0.1 N, biases ~ 0.02 N,
relative-position tables ~ 0.5 N; all branches (incl. cva.crossattn.proj_out) are live and the
deformable softmax is not saturated (SURVEY section 8d "W2").
"""
import zlib

import torch

_BUFFER_SUFFIXES = ("relative_position_index", "attn_mask")


def is_buffer_key(key: str) -> bool:
    return key.endswith(_BUFFER_SUFFIXES)


def seeded_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if key.endswith("relative_position_bias_table"):
        return 0.5 * r
    if key == "final_out.bias":
        return 0.02 * r + 0.52                # centres the synthetic logits on the mask threshold (worst case for a20)
    if key.endswith(".bias"):
        return 0.02 * r
    if len(shape) == 1:                       # LayerNorm / GroupNorm scale
        return 1.0 + 0.1 * r
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 1.0
    if ".cva.crossattn.conv_offset.3." in key:
        gain = 2.0                            # make the sampling offsets non-trivial
    return r * (gain / fan_in ** 0.5)


def fill_state_dict(template: dict, seed: int = 0) -> dict:
    """template: key -> tensor (shape/dtype source; integer/buffer entries are kept as they are)."""
    out = {}
    for k, v in template.items():
        if is_buffer_key(k) or not torch.is_floating_point(v):
            out[k] = v.clone()
        else:
            out[k] = seeded_tensor(k, v.shape, seed)
    return out


def from_manifest(manifest: dict, seed: int = 0) -> dict:
    """manifest: key -> list(shape) for the float parameters only (buffers are rebuilt by the modules)."""
    return {k: seeded_tensor(k, shp, seed) for k, shp in manifest.items()}
