"""ControlLoRA restated (oracle; test infrastructure).

Follows /root/reference/model/controllora.py:
  * ControlLoRAModel.__init__ LoRA injection             :529-593  (`_skip_layers` :443-450)
  * tie_weights / _tie_weights                           :45-56, :623-632
  * state_dict / load_state_dict filter semantics         :600-614
  * fuse_lora / fuse                                     :728-777
  * VAEControlNetConditioningEmbedding (conv_vae_out aliases conv_in) :28-42, :596-598
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
from torch import nn

from .sd15 import (
    ControlNetModel,
    LoRACompatibleConv,
    LoRACompatibleLinear,
    LoRAConv2dLayer,
    LoRALinearLayer,
    SD15Config,
    UNet2DConditionModel,
)

_SKIP_LAYERS = ["conv_in", "time_proj", "time_embedding", "class_embedding", "down_blocks", "mid_block"]


def _tie_weights(source: nn.Module, target: nn.Module) -> None:
    """Re-point every Parameter of `target` at the same-named Parameter of `source` (:45-56)."""
    for name, _ in list(source.named_parameters()):
        *branches, base = name.split(".")
        s, t = source, target
        for b in branches:
            s, t = getattr(s, b), getattr(t, b)
        setattr(t, base, getattr(s, base))


class VAEControlNetConditioningEmbedding(nn.Module):
    """embedding = conv_vae_out(vae.encode(img).sample() * scaling_factor); conv_vae_out IS conv_in (:36)."""

    def __init__(self, conv_unet: nn.Conv2d, autoencoder=None, scaling_factor: float = 0.18215):
        super().__init__()
        self.autoencoder = autoencoder
        self.scaling_factor = scaling_factor
        self.conv_vae_out = conv_unet
        for p in self.conv_vae_out.parameters():  # zero_module (:36)
            nn.init.zeros_(p)

    def forward(self, conditioning):
        emb = self.autoencoder.encode(conditioning).latent_dist.sample() * self.scaling_factor
        return self.conv_vae_out(emb)


class ControlLoRAModel(ControlNetModel):
    _skip_layers = _SKIP_LAYERS

    def __init__(
        self,
        cfg: Optional[SD15Config] = None,
        lora_linear_rank: int = 4,
        lora_conv2d_rank: int = 0,
        uses_vae: bool = True,
    ):
        super().__init__(cfg)
        self.lora_linear_rank = lora_linear_rank
        self.lora_conv2d_rank = lora_conv2d_rank
        for name, layer in list(self.named_modules()):
            if name.split(".")[0] not in self._skip_layers:
                continue
            if lora_conv2d_rank > 0 and isinstance(layer, LoRACompatibleConv):
                # sic: rank=lora_linear_rank (controllora.py:569)
                layer.set_lora_layer(
                    LoRAConv2dLayer(
                        layer.in_channels,
                        layer.out_channels,
                        rank=lora_linear_rank,
                        kernel_size=layer.kernel_size,
                        stride=layer.stride,
                        padding=layer.padding,
                    )
                )
            elif lora_linear_rank > 0 and isinstance(layer, LoRACompatibleLinear):
                layer.set_lora_layer(LoRALinearLayer(layer.in_features, layer.out_features, rank=lora_linear_rank))
        if uses_vae:
            self.controlnet_cond_embedding = VAEControlNetConditioningEmbedding(conv_unet=self.conv_in)

    # -- reference surface ---------------------------------------------------------------
    @classmethod
    def from_unet(cls, unet: UNet2DConditionModel, lora_linear_rank: int = 4, lora_conv2d_rank: int = 0,
                  autoencoder=None):
        net = cls(unet.cfg, lora_linear_rank=lora_linear_rank, lora_conv2d_rank=lora_conv2d_rank,
                  uses_vae=autoencoder is not None or True)
        net.tie_weights(unet)
        if autoencoder is not None:
            net.set_autoencoder(autoencoder)
        return net

    def tie_weights(self, unet: UNet2DConditionModel):  # :623-632
        _tie_weights(unet.conv_in, self.conv_in)
        _tie_weights(unet.time_embedding, self.time_embedding)
        _tie_weights(unet.down_blocks, self.down_blocks)
        _tie_weights(unet.mid_block, self.mid_block)

    def set_autoencoder(self, autoencoder):
        if isinstance(self.controlnet_cond_embedding, VAEControlNetConditioningEmbedding):
            self.controlnet_cond_embedding.autoencoder = autoencoder
        else:
            self.controlnet_cond_embedding = VAEControlNetConditioningEmbedding(self.conv_in, autoencoder)

    def state_dict(self, *args, **kwargs):  # :600-606
        sd = super().state_dict(*args, **kwargs)
        return OrderedDict(
            (k, v) for k, v in sd.items() if k.split(".")[0] not in self._skip_layers or ".lora_layer." in k
        )

    def full_state_dict(self):
        return super().state_dict()

    def load_state_dict(self, state_dict, strict: bool = True):  # :608-614
        new = OrderedDict(state_dict)
        for k, v in super().state_dict().items():
            if k.split(".")[0] in self._skip_layers and k not in new:
                new[k] = v
        return super().load_state_dict(new, strict)

    def fuse_lora(self, lora_scale: float = 1.0):  # :728-737
        for m in self.modules():
            if isinstance(m, (LoRACompatibleConv, LoRACompatibleLinear)):
                m._fuse_lora(lora_scale)

    def lora_param_count(self) -> int:
        return sum(p.numel() for n, p in self.named_parameters() if ".lora_layer." in n)
