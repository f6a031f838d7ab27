"""Model factory mirroring reference models/factory/modelFactory.py: create_view_config (:17-33) and
create_multiswin (:36-73) with the same hard-coded hyper-parameters.  ml_collections is not a dependency:
ConfigDict below gives the attribute/item access the reference code uses (`cfg["k"]`, `cfg.patches.size`).
"""
import os

import torch

from ..encoder.multiTemporalViewEncoder import ThreeViewSwinTransformer


class ConfigDict(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def load_model_weights(model, path, strict=False):
    state_dict = torch.load(path, map_location="cpu")
    model.load_state_dict(state_dict, strict=strict)
    return model


def create_view_config(hidden_sizes, patches_size, depths, num_heads, mlp_dim, num_frames, input_resolution, temporal_dim,
                       temporal_ratio=None, window_size=7):
    """The reference hard-codes window_size 7 (:25); the extra keyword serves the patched-resolution configurations
    (512x512 needs window 8)."""
    return ConfigDict({
        'hidden_size': hidden_sizes,
        'patches': {'size': patches_size},
        'window_size': window_size,
        'depths': depths,
        'num_heads': num_heads,
        'mlp_dim': mlp_dim,
        'num_frames': num_frames,
        'input_resolution': input_resolution,
        'temporal_dim': temporal_dim,
        'temporal_ratio': temporal_ratio or [1] * len(depths),
    })


def default_view_configs(img_size=224, window_size=7):
    """modelFactory.py:38-45 at 224 / window 7; other sizes follow SURVEY A10 (token grid img/4 .. img/32 must be divisible by
    the window where it is larger than it)."""
    res = [(img_size // d, img_size // d) for d in (4, 8, 16, 32)]
    for r, _ in res:
        if img_size % 32 or (r > window_size and r % window_size):
            raise ValueError("image size %d does not tile with window %d (stage grid %d)" % (img_size, window_size, r))
    kw = dict(window_size=window_size)
    return [
        create_view_config([96, 192, 384, 768], (4, 4, 3), [2, 2, 6, 2], [3, 6, 12, 24], 768, 1, res, 1, [1, 1], **kw),
        create_view_config([96, 192, 384, 768], (4, 4, 2), [2, 2, 18, 2], [3, 6, 12, 24], 1536, 1, res, 1, [1, 3], **kw),
        create_view_config([128, 256, 512, 1024], (4, 4, 1), [2, 2, 18, 2], [4, 8, 16, 32], 3072, 3, res, 3, **kw),
    ]


def create_multiswin(weights_path="../weights/weight.pth", img_size=224, window_size=7):
    """Same model as the reference factory.  The reference unconditionally torch.load()s ../weights/weight.pth
    (strict=False, :70-71); here the ImageNet-style init file is loaded when it exists and skipped otherwise, so the
    model can be built for random-init / checkpoint-restore use without it."""
    view_configs = default_view_configs(img_size, window_size)
    global_encoder_config = ConfigDict({'num_heads': 12, 'mlp_dim': 3072, 'num_layers': 12, 'hidden_size': 768,
                                        'merge_axis': 'channel', 'num_frames': 3})
    model = ThreeViewSwinTransformer(view_configs=view_configs, input_token_temporal_dims=[1, 1, 3],
                                     global_encoder_config=global_encoder_config)
    if weights_path and os.path.exists(weights_path):
        model = load_model_weights(model, weights_path, strict=False)
    return model, view_configs
