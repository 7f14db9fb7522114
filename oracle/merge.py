"""EdgeStyle multi-ControlNet merge restated (oracle; test infrastructure).

Follows /root/reference/model/edgestyle_multicontrolnet.py:
  * ControlNetBlock                         :23-63
  * EdgeStyleMultiControlNetModel.forward   :116-171 (sequential nets :133-152, interleave :160-164,
                                                      13 merge blocks :167-169; always returns a tuple)
  * interleave_tensors                      :479-514
  * down_output_channels / down_sizes table :73-102 -- parameterised here from (block_out_channels,
    latent h, w) because the reference hard-codes 512x512.
`closed_form_block` is SURVEY.md A.9 (what the CUDA merge kernel implements).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .sd15 import SD15Config


class ControlNetBlock(nn.Module):
    def __init__(self, output_channel: int, size: Tuple[int, int], num_controlnets: int):
        super().__init__()
        half = output_channel * num_controlnets // 2
        self.first_conv = nn.Conv2d(output_channel * num_controlnets, half, kernel_size=1, groups=half)
        self.first_normalization = nn.LayerNorm([half, *size])
        self.activation = nn.SiLU()
        self.second_conv = nn.Conv2d(half, output_channel, kernel_size=1, groups=output_channel)
        self.second_normalization = nn.LayerNorm([output_channel, *size])
        self.third_conv = nn.Conv2d(output_channel, output_channel, kernel_size=1, groups=output_channel)

    def forward(self, x):
        x = self.activation(self.first_normalization(self.first_conv(x)))
        x = self.activation(self.second_normalization(self.second_conv(x)))
        return self.third_conv(x)


def interleave_tensors(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    assert all(t.size() == tensors[0].size() for t in tensors)
    stacked = torch.stack(list(tensors), dim=1)  # [B, n, C, H, W]
    b, n, c, h, w = stacked.shape
    return stacked.permute(0, 2, 1, 3, 4).contiguous().view(b, -1, h, w)


def residual_shapes(cfg: SD15Config, h: int, w: int) -> List[Tuple[int, int, int]]:
    """(C, H, W) of the 12 down residuals + mid, for a latent of h x w (stride-2 convs: ceil)."""
    boc = cfg.block_out_channels
    shapes = [(boc[0], h, w)]
    for i, c in enumerate(boc):
        shapes += [(c, h, w)] * cfg.layers_per_block
        if i != len(boc) - 1:
            h, w = (h + 1) // 2, (w + 1) // 2
            shapes.append((c, h, w))
    shapes.append((boc[-1], h, w))
    return shapes


class EdgeStyleMultiControlNetModel(nn.Module):
    def __init__(self, controlnets: Sequence[nn.Module], cfg: SD15Config = None, latent_hw: Tuple[int, int] = (64, 64)):
        super().__init__()
        self.nets = nn.ModuleList(controlnets)
        cfg = cfg or controlnets[0].cfg
        shapes = residual_shapes(cfg, *latent_hw)
        n = len(controlnets)
        self.multi_controlnet_down_blocks = nn.ModuleList([ControlNetBlock(c, (h, w), n) for c, h, w in shapes[:-1]])
        c, h, w = shapes[-1]
        self.multi_controlnet_mid_block = ControlNetBlock(c, (h, w), n)

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale,
                guess_mode: bool = False, return_dict: bool = True):
        downs, mids = [], []
        for image, scale, net in zip(controlnet_cond, conditioning_scale, self.nets):
            d, m = net(sample, timestep, encoder_hidden_states, image, scale, guess_mode=guess_mode)
            downs.append(d)
            mids.append(m)
        down = [interleave_tensors(ts) for ts in zip(*downs)]
        mid = interleave_tensors(mids)
        down = [blk(x) for blk, x in zip(self.multi_controlnet_down_blocks, down)]
        mid = self.multi_controlnet_mid_block(mid)
        return down, mid

    def merge_state_dict(self):
        """Only the merge blocks (edgestyle_multicontrolnet.py:173-193)."""
        return {k: v for k, v in self.state_dict().items() if k.startswith("multi_controlnet_")}


def closed_form_block(block: ControlNetBlock, residuals: Sequence[torch.Tensor]) -> torch.Tensor:
    """the merge without materialising the interleaved tensor.

    residuals: n tensors [B,C,H,W] (already x conditioning_scale).  Net 2p pairs with net 2p+1.
    """
    n = len(residuals)
    b, c, h, w = residuals[0].shape
    P = n // 2
    w1 = block.first_conv.weight.view(c, P, 2)  # group g = c*P + p, inputs (c*n + 2p, c*n + 2p + 1)
    b1 = block.first_conv.bias.view(c, P)
    u = torch.stack(
        [
            w1[:, p, 0].view(1, c, 1, 1) * residuals[2 * p]
            + w1[:, p, 1].view(1, c, 1, 1) * residuals[2 * p + 1]
            + b1[:, p].view(1, c, 1, 1)
            for p in range(P)
        ],
        dim=2,
    )  # [B, C, P, H, W]
    mu = u.mean(dim=(1, 2, 3, 4), keepdim=True)
    var = u.var(dim=(1, 2, 3, 4), unbiased=False, keepdim=True)
    g1 = block.first_normalization.weight.view(1, c, P, h, w)
    be1 = block.first_normalization.bias.view(1, c, P, h, w)
    v = F.silu((u - mu) / torch.sqrt(var + 1e-5) * g1 + be1)
    w2 = block.second_conv.weight.view(c, P)
    z = (v * w2.view(1, c, P, 1, 1)).sum(dim=2) + block.second_conv.bias.view(1, c, 1, 1)
    mu2 = z.mean(dim=(1, 2, 3), keepdim=True)
    var2 = z.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    y = F.silu((z - mu2) / torch.sqrt(var2 + 1e-5) * block.second_normalization.weight[None]
               + block.second_normalization.bias[None])
    return y * block.third_conv.weight.view(1, c, 1, 1) + block.third_conv.bias.view(1, c, 1, 1)
