"""Encoder wrapper mirroring reference models/encoder/encoder.py:6-18 (state_dict keys under `base.`)."""
from torch import nn

from ..factory.modelFactory import create_multiswin


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.base, self.configs = create_multiswin()

    def forward(self, x, return_attention=False, layer_id=1):
        """x (B,3,3,224,224) -> (final_x (B,2304,7,7), view_x [4 stages][3 views] of (B,1,L,C), dct_x (B,9,224,224)).

        final_x is the (B,49,2304) token matrix viewed as b c h w (a channels-last view: no transpose kernel runs;
        the reference's einops rearrange returns the same kind of view)."""
        ws = self.configs[0]["window_size"]
        final_x, view_x, dct_x = self.base(x)
        if not return_attention:
            B, n, C = final_x.shape
            side = int(round(n ** 0.5))
            assert side * side == n and side == ws
            final_x = final_x.view(B, side, side, C).permute(0, 3, 1, 2)
        return final_x, view_x, dct_x
