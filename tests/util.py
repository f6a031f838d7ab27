"""Shared helpers for the parity tests: golden fixtures, seeded inputs, key-seeded weights."""
import json
import os

import torch

from oracle import weights as wts

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seeded_input(shape, seed):
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def golden(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def manifest():
    with open(os.path.join(GOLD, "manifest.json")) as f:
        return json.load(f)


def seeded_state_dict(module, seed=0):
    """Key-seeded weights for every float parameter of `module` (buffers are kept)."""
    return wts.fill_state_dict({k: v.detach().cpu() for k, v in module.state_dict().items()}, seed)


def load_seeded(module, seed=0):
    sd = seeded_state_dict(module, seed)
    module.load_state_dict(sd, strict=True)
    return sd


def maxabs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def decoder_inputs(B=1):
    from oracle import mumpy_oracle as orc
    final_x = seeded_input((B, 2304, 7, 7), 81)
    ff = seeded_input((B, 9, 224, 224), 82)
    view_x = []
    for s in range(4):
        h = (56, 28, 14, 7)[s]
        view_x.append([seeded_input((B, 1, orc.VIEW_T[v] * h * h, orc.VIEW_DIMS[v][s]), 83 + 3 * s + v) for v in range(3)])
    return final_x, view_x, ff
