"""Step engine: weight packing + static launch schedule of the denoise step (filled in below)."""
from __future__ import annotations

from typing import Dict

import torch


def pack_merge_block(sd: Dict[str, torch.Tensor], C: int, h: int, w: int, dtype, device) -> Dict[str, torch.Tensor]:
    """Repack one ControlNetBlock (/root/reference/model/edgestyle_multicontrolnet.py:23-63) for es_merge_phase.

    The reference's interleaved channel index is c*6 + net; first_conv group g = c*3 + p consumes nets
    (2p, 2p+1) of channel c.  LayerNorm affine [3C, H, W] / [C, H, W] become channels-last [hw, 3, C] / [hw, C].
    """
    f32 = dict(device=device, dtype=torch.float32)
    hw = h * w
    out = {
        "w1": sd["first_conv.weight"].reshape(C, 3, 2).to(**f32).contiguous(),
        "b1": sd["first_conv.bias"].reshape(C, 3).to(**f32).contiguous(),
        "g1": sd["first_normalization.weight"].reshape(C, 3, hw).permute(2, 1, 0).to(device=device, dtype=dtype).contiguous(),
        "be1": sd["first_normalization.bias"].reshape(C, 3, hw).permute(2, 1, 0).to(device=device, dtype=dtype).contiguous(),
        "w2": sd["second_conv.weight"].reshape(C, 3).to(**f32).contiguous(),
        "b2": sd["second_conv.bias"].reshape(C).to(**f32).contiguous(),
        "g2": sd["second_normalization.weight"].reshape(C, hw).permute(1, 0).to(device=device, dtype=dtype).contiguous(),
        "be2": sd["second_normalization.bias"].reshape(C, hw).permute(1, 0).to(device=device, dtype=dtype).contiguous(),
        "w3": sd["third_conv.weight"].reshape(C).to(**f32).contiguous(),
        "b3": sd["third_conv.bias"].reshape(C).to(**f32).contiguous(),
    }
    return out
