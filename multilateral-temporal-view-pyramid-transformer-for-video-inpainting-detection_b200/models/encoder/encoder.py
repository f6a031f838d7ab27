"""Encoder wrapper mirroring reference models/encoder/encoder.py:6-18 (state_dict keys under `base.`)."""
from torch import nn

from ..factory.modelFactory import create_multiswin


class Encoder(nn.Module):
    def __init__(self, img_size=224, window_size=7):
        """Zero-argument construction is the reference's (224x224, window 7).  img_size / window_size select the
        patched-resolution configurations of SURVEY A10 (e.g. 512, 8 -- BASELINE.json configs[3]); pair them with
        Decoder(shape=[img_size // 4, img_size // 8, img_size // 16, img_size // 32])."""
        super().__init__()
        self.base, self.configs = create_multiswin(img_size=img_size, window_size=window_size)

    def forward(self, x, return_attention=False, layer_id=1):
        """x (B,3,3,224,224) -> (final_x (B,2304,7,7), view_x [4 stages][3 views] of (B,1,L,C), dct_x (B,9,224,224)).

        final_x is the (B,49,2304) token matrix viewed as b c h w (a channels-last view: no transpose kernel runs;
        the reference's einops rearrange returns the same kind of view).  The reference writes h = w = window_size
        (encoder.py:16-17), which equals the last stage's grid side only at 224x224; the grid side is used here."""
        final_x, view_x, dct_x = self.base(x)
        if not return_attention:
            B, n, C = final_x.shape
            side = int(round(n ** 0.5))
            assert side * side == n
            final_x = final_x.view(B, side, side, C).permute(0, 3, 1, 2)
        return final_x, view_x, dct_x
