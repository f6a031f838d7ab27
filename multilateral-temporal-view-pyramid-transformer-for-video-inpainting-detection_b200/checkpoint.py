"""Checkpoint ingestion: the authors' result folders hold `encoder_{epoch}.pt` /
`decoder_{epoch}.pt` state_dicts written by utils/utils.py:286-297 (plus optimiser states and args.pkl that inference does
not need).  `load_checkpoint` reads the two model files the way utils/utils.py:301-321 does, `strip_data_parallel` is
check_parallel (utils/utils.py:156-176: the `module.` prefix of nn.DataParallel is removed from both dicts when the FIRST
encoder key carries it), and `restore` loads them strictly into the mirror modules -- whose state_dict keys are the
reference's -- and pre-packs the kernel-side operand copies by running nothing more than a load (packing is lazy and keyed
on the parameter versions, see models/modules/_packing.py).
"""
import os
from collections import OrderedDict

import torch


def strip_data_parallel(encoder_dict, decoder_dict):
    """utils/utils.py:156-176, same decision rule (only the first encoder key is inspected)."""
    trained_parallel = False
    for k in encoder_dict:
        if k[:7] == "module.":
            trained_parallel = True
        break
    if not trained_parallel:
        return encoder_dict, decoder_dict
    return (OrderedDict((k[7:], v) for k, v in encoder_dict.items()),
            OrderedDict((k[7:], v) for k, v in decoder_dict.items()))


def load_checkpoint(model_dir, epoch=3, map_location="cpu"):
    """(encoder_dict, decoder_dict) from `<model_dir>/encoder_<epoch>.pt` and `decoder_<epoch>.pt` (utils/utils.py:301-321 reads
    `../results/<model_name>/…`; pass that folder).  Optimiser states and args.pkl are not touched."""
    enc = torch.load(os.path.join(model_dir, "encoder_%s.pt" % epoch), map_location=map_location, weights_only=True)
    dec = torch.load(os.path.join(model_dir, "decoder_%s.pt" % epoch), map_location=map_location, weights_only=True)
    return strip_data_parallel(enc, dec)


def restore(encoder, decoder, model_dir, epoch=3, device=None):
    """Loads a reference checkpoint into mumpy_b200 Encoder/Decoder (strict, like test.py:60-61), moves them to `device`
    and puts them in eval mode."""
    enc_sd, dec_sd = load_checkpoint(model_dir, epoch)
    encoder.load_state_dict(enc_sd, strict=True)
    decoder.load_state_dict(dec_sd, strict=True)
    from .models.modules._packing import invalidate_packed
    invalidate_packed(encoder)
    invalidate_packed(decoder)
    if device is not None:
        encoder.to(device)
        decoder.to(device)
    return encoder.eval(), decoder.eval()
