"""AutoencoderKL (SD1.5 VAE) restated in plain PyTorch fp32 -- ORACLE, test infrastructure only.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU baseline may import this module; the product path
(`edgestyle_b200/vae.py`) never does.

What it restates: `diffusers==0.26.3` `AutoencoderKL` (models/autoencoders/autoencoder_kl.py, vae.py;
`Encoder`, `Decoder`, `DownEncoderBlock2D`, `UpDecoderBlock2D`, `UNetMidBlock2D` with the single-head
`Attention(_from_deprecated_attn_block=True)`, `DiagonalGaussianDistribution`).  The reference reaches it at
/root/reference/model/controllora.py:38-42 (`autoencoder.encode(conditioning).latent_dist.sample()` times
`config.scaling_factor`, once per call through edgestyle_pipeline.py:660-662) and at
/root/reference/model/edgestyle_pipeline.py:552-557 (`vae.decode(latents / scaling_factor)`).  diffusers is a
third-party dependency absent from /root/reference and from this image, and the reference holds no golden vectors
for this stage: **parity unpinned** (the restatement is anchored on the published layer list and the
diffusers state-dict key names, which `state_dict()` reproduces so that a real `vae/` checkpoint loads unchanged:
SD1.5 VAE = 83 653 863 parameters, checked in tests/test_oracle_golden.py and tests/test_vae_host_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn


@dataclass
class VaeConfig:
    """SD1.5 `vae/config.json`; shrinkable for fast tests."""

    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    norm_eps: float = 1e-6


class VaeResnet(nn.Module):
    """ResnetBlock2D(temb_channels=None, output_scale_factor=1): x + conv2(silu(gn(conv1(silu(gn(x))))))."""

    def __init__(self, cin: int, cout: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):
    """Attention(heads=1, dim_head=C, bias=True, residual_connection=True, norm_num_groups, eps) over h*w tokens."""

    def __init__(self, c: int, groups: int, eps: float):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        a = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1) @ v
        a = self.to_out[0](a)
        return x + a.transpose(1, 2).reshape(b, c, h, w)


class _Conv(nn.Module):
    def __init__(self, c: int, stride: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=0 if stride == 2 else 1)


class VaeDownsample(_Conv):
    """Downsample2D(use_conv=True, padding=0): F.pad(x, (0, 1, 0, 1)) then a stride-2 conv."""

    def __init__(self, c: int):
        super().__init__(c, 2)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class VaeUpsample(_Conv):
    """Upsample2D(use_conv=True): nearest x2 then a 3x3 conv."""

    def __init__(self, c: int):
        super().__init__(c, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self, resnets, sampler, key):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if sampler is not None:
            setattr(self, key, nn.ModuleList([sampler]))
        self._key = key

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        s = getattr(self, self._key, None)
        return s[0](x) if s is not None else x


class VaeMid(nn.Module):
    def __init__(self, c, groups, eps):
        super().__init__()
        self.attentions = nn.ModuleList([VaeAttention(c, groups, eps)])
        self.resnets = nn.ModuleList([VaeResnet(c, c, groups, eps), VaeResnet(c, c, groups, eps)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, cfg: VaeConfig):
        super().__init__()
        ch, g, e = cfg.block_out_channels, cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        blocks, cin = [], ch[0]
        for i, cout in enumerate(ch):
            res = [VaeResnet(cin if j == 0 else cout, cout, g, e) for j in range(cfg.layers_per_block)]
            blocks.append(_Block(res, VaeDownsample(cout) if i < len(ch) - 1 else None, "downsamplers"))
            cin = cout
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = VaeMid(ch[-1], g, e)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=e)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg: VaeConfig):
        super().__init__()
        ch, g, e = cfg.block_out_channels, cfg.norm_num_groups, cfg.norm_eps
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, padding=1)
        self.mid_block = VaeMid(rev[0], g, e)
        blocks, cin = [], rev[0]
        for i, cout in enumerate(rev):
            res = [VaeResnet(cin if j == 0 else cout, cout, g, e) for j in range(cfg.layers_per_block + 1)]
            blocks.append(_Block(res, VaeUpsample(cout) if i < len(ch) - 1 else None, "upsamplers"))
            cin = cout
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=e)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None):
        if noise is None:  # diffusers: randn_tensor(mean.shape, generator=generator, device=..., dtype=...)
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: Optional[VaeConfig] = None):
        super().__init__()
        self.cfg = self.config = cfg or VaeConfig()
        self.encoder = Encoder(self.cfg)
        self.decoder = Decoder(self.cfg)
        self.quant_conv = nn.Conv2d(2 * self.cfg.latent_channels, 2 * self.cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(self.cfg.latent_channels, self.cfg.latent_channels, 1)

    def encode(self, x):
        return _Out(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return _Out(sample=self.decoder(self.post_quant_conv(z)))
