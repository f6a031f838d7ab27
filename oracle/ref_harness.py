"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference on CPU.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to pin the oracle restatement and to generate tests/golden/*.

Work-arounds, none of which touch reference arithmetic:
  * `timm.models.layers` / `ml_collections` are not installed -> tiny shim modules
    (DropPath is identity in eval, trunc_normal_ = torch.nn.init.trunc_normal_).
  * models/modules/dct.py:16,18,61,62 call `.cuda()` in constructors -> Tensor.cuda no-op on CPU.
  * models/factory/modelFactory.py:70-71 torch.load("../weights/weight.pth") is bypassed by
    building ThreeViewSwinTransformer with the constants of modelFactory.py:38-62.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MUMPY_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


class _ConfigDict(dict):
    """ml_collections.ConfigDict stand-in: dict with attribute access, nested dicts converted."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = _ConfigDict(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training, "shim DropPath is eval-only"
        return x


def _install_shims():
    if "timm.models.layers" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers
    if "ml_collections" not in sys.modules:
        mlc = types.ModuleType("ml_collections")
        mlc.ConfigDict = _ConfigDict
        sys.modules["ml_collections"] = mlc
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


_ref_modules = None


def load():
    """Returns a namespace with the reference's hot-path modules."""
    global _ref_modules
    if _ref_modules is not None:
        return _ref_modules
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_shims()
    # the reference uses top-level package name `models`; make sure ours does not shadow it
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import models.encoder.multiTemporalViewEncoder as mtv
        import models.modules.swinTransformer as swin
        import models.modules.blocks as blocks
        import models.modules.deformableAttention as datt
        import models.modules.dct as dct
        import models.decoder.decoder as dec
        import models.factory.modelFactory as fac
    finally:
        sys.path.remove(REFERENCE_ROOT)
    ns = types.SimpleNamespace(mtv=mtv, swin=swin, blocks=blocks, datt=datt, dct=dct, dec=dec, fac=fac)
    _ref_modules = ns
    return ns


def view_configs(ref, window_size=7, res=(56, 28, 14, 7)):
    """modelFactory.py:38-45 constants (optionally patched resolution/window)."""
    r = [(s, s) for s in res]
    cfgs = [
        ref.fac.create_view_config([96, 192, 384, 768], (4, 4, 3), [2, 2, 6, 2], [3, 6, 12, 24], 768, 1, r, 1, [1, 1]),
        ref.fac.create_view_config([96, 192, 384, 768], (4, 4, 2), [2, 2, 18, 2], [3, 6, 12, 24], 1536, 1, r, 1, [1, 3]),
        ref.fac.create_view_config([128, 256, 512, 1024], (4, 4, 1), [2, 2, 18, 2], [4, 8, 16, 32], 3072, 3, r, 3),
    ]
    for c in cfgs:
        c["window_size"] = window_size
    return cfgs


def build_reference_encoder(depths=(2, 2, 18, 2)):
    """ThreeViewSwinTransformer exactly as create_multiswin() builds it, minus the torch.load."""
    ref = load()
    cfgs = view_configs(ref)
    genc = _ConfigDict({'num_heads': 12, 'mlp_dim': 3072, 'num_layers': 12, 'hidden_size': 768,
                        'merge_axis': 'channel', 'num_frames': 3})
    model = ref.mtv.ThreeViewSwinTransformer(view_configs=cfgs, input_token_temporal_dims=[1, 1, 3],
                                             global_encoder_config=genc, depths=list(depths))
    return model.eval(), cfgs


class ReferenceEncoder(nn.Module):
    """models/encoder/encoder.py:6-18 with the factory's torch.load bypassed (state_dict prefix `base.`)."""

    def __init__(self):
        super().__init__()
        self.base, self.configs = build_reference_encoder()

    def forward(self, x):
        from einops import rearrange
        ws = self.configs[0]["window_size"]
        out_channels = self.configs[1]["hidden_size"][-1] * 3
        final_x, view_x, dct_x = self.base(x)
        final_x = rearrange(final_x, "b (h w) (p1 p2 c) -> b c (h p1) (w p2)", p1=1, p2=1, h=ws, w=ws, c=out_channels)
        return final_x, view_x, dct_x


def build_reference_decoder():
    ref = load()
    return ref.dec.Decoder().eval()
