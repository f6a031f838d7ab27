"""CPU: drop-in boundary -- the mirror classes expose exactly the reference's state_dict keys and shapes
(manifest captured from the reference by oracle/make_golden.py), buffers included."""
import torch

from tests import util


def test_encoder_decoder_keys_match_reference():
    import mumpy_b200
    m = util.manifest()
    enc, dec = mumpy_b200.Encoder(), mumpy_b200.Decoder()
    e = {k: list(v.shape) for k, v in enc.state_dict().items()}
    d = {k: list(v.shape) for k, v in dec.state_dict().items()}
    ref_e = dict(m["encoder"])
    ref_e.update(m["encoder_buffers"])
    assert e == ref_e and len(e) == 1172
    assert d == m["decoder"] and len(d) == 92


def test_buffers_equal_oracle_definitions():
    from oracle import mumpy_oracle as orc
    from mumpy_b200.models.modules.swinTransformer import SwinTransformerBlock
    blk = SwinTransformerBlock(64, (14, 14), 2, window_size=7, shift_size=3, temporal_dim=3)
    assert torch.equal(blk.attn_mask, orc.shifted_window_mask(42, 14, 7, 3))
    table = torch.arange(169 * 2, dtype=torch.float32).view(169, 2)
    assert torch.equal(table[blk.attn.relative_position_index.view(-1)].view(49, 49, 2).permute(2, 0, 1),
                       orc.relative_position_bias(table, 7))


def test_constructor_signatures():
    import inspect
    import mumpy_b200
    from mumpy_b200.models.modules import swinTransformer as sw, blocks, deformableAttention as da
    from mumpy_b200.models.encoder import multiTemporalViewEncoder as mtv
    assert list(inspect.signature(mumpy_b200.Decoder.__init__).parameters)[1:] == [
        "in_channels", "out_channels", "kernel_size", "num_classes", "dap_k", "features", "input_token_temporal_dims",
        "rgb_features", "shape"]
    assert list(inspect.signature(sw.SwinTransformerBlock.__init__).parameters)[1:] == [
        "dim", "input_resolution", "num_heads", "window_size", "shift_size", "mlp_ratio", "qkv_bias", "qk_scale", "drop",
        "attn_drop", "drop_path", "act_layer", "norm_layer", "temporal_dim", "fused_window_process"]
    assert list(inspect.signature(da.SwinDAttention.__init__).parameters)[1:] == [
        "dim1", "n_heads", "attn_drop", "n_groups", "ws", "stride", "offset_range_factor", "no_off", "height_scale",
        "dwc_pe", "use_pe", "fixed_pe"]
    assert list(inspect.signature(blocks.Block.__init__).parameters)[1:] == ["dim", "heads", "mlp_dim", "dropout", "drop_path"]
    assert list(inspect.signature(mtv.ThreeViewSwinTransformer.__init__).parameters)[1:4] == [
        "view_configs", "input_token_temporal_dims", "global_encoder_config"]


def test_checkpoint_ingestion_roundtrip(tmp_path):
    """utils/utils.py:286-321 + check_parallel (:156-176): a checkpoint written from DataParallel-wrapped reference modules
    (keys prefixed `module.`) and a plain one both restore strictly into the mirrors."""
    import mumpy_b200
    from mumpy_b200 import checkpoint
    enc_sd = util.seeded_state_dict(mumpy_b200.Encoder())       # every key of the reference state_dict, buffers included
    dec_sd = util.seeded_state_dict(mumpy_b200.Decoder())
    for prefix in ("", "module."):
        d = tmp_path / ("ckpt_" + (prefix.strip(".") or "plain"))
        d.mkdir()
        torch.save({prefix + k: v for k, v in enc_sd.items()}, d / "encoder_3.pt")
        torch.save({prefix + k: v for k, v in dec_sd.items()}, d / "decoder_3.pt")
        e, dd = checkpoint.load_checkpoint(str(d), 3)
        assert list(e.keys()) == list(enc_sd.keys()) and list(dd.keys()) == list(dec_sd.keys())
        assert all(torch.equal(e[k], enc_sd[k]) for k in enc_sd)
    enc, dec = mumpy_b200.Encoder(), mumpy_b200.Decoder()
    enc, dec = checkpoint.restore(enc, dec, str(d), 3)
    assert not enc.training and not dec.training
    got = enc.state_dict()
    assert all(torch.equal(got[k], enc_sd[k]) for k in enc_sd)
