"""SD1.5 UNet2DConditionModel / ControlNetModel restated in plain PyTorch (oracle; test infrastructure).

Follows diffusers==0.26.3 semantics as summarised in SURVEY.md Appendix A.0-A.4 (the reference
imports these classes at /root/reference/model/controllora.py:8-22 and
/root/reference/model/edgestyle_pipeline.py:12-54; their source is not in the tree).  Module and
parameter names reproduce the diffusers state-dict keys (Appendix A.7) so that real checkpoints
load unchanged.  Every ``nn.Linear`` is a :class:`LoRACompatibleLinear` and every conv a
:class:`LoRACompatibleConv` whose ``lora_layer`` is ``None`` unless ControlLoRA injects one
(/root/reference/model/controllora.py:529-593).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn


@dataclass
class SD15Config:
    """SD1.5 `unet/config.json`; shrinkable for fast CPU tests."""

    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    num_heads: int = 8  # diffusers' `attention_head_dim=8` quirk: 8 heads at every level
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    conditioning_embedding_out_channels: Tuple[int, ...] = (16, 32, 96, 256)
    conditioning_channels: int = 3
    # which down blocks carry transformers (CrossAttnDownBlock2D x3, DownBlock2D)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


# --------------------------------------------------------------------------------------------
# LoRA-compatible primitives (diffusers models/lora.py semantics)
# --------------------------------------------------------------------------------------------
class LoRALinearLayer(nn.Module):
    def __init__(self, in_features: int, out_features: int, rank: int = 4):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        self.rank = rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, x):
        return self.up(self.down(x.to(self.down.weight.dtype))).to(x.dtype)


class LoRAConv2dLayer(nn.Module):
    def __init__(self, in_features, out_features, rank=4, kernel_size=(1, 1), stride=(1, 1), padding=0):
        super().__init__()
        self.down = nn.Conv2d(in_features, rank, kernel_size, stride, padding, bias=False)
        self.up = nn.Conv2d(rank, out_features, (1, 1), (1, 1), bias=False)
        self.rank = rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, x):
        return self.up(self.down(x.to(self.down.weight.dtype))).to(x.dtype)


class LoRACompatibleLinear(nn.Linear):
    """y = W x + b + scale * up(down(x)); `_fuse_lora` folds W += scale * up @ down."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer: Optional[LoRALinearLayer] = None

    def set_lora_layer(self, lora_layer):
        self.lora_layer = lora_layer

    def _fuse_lora(self, lora_scale: float = 1.0):
        if self.lora_layer is None:
            return
        w_up = self.lora_layer.up.weight.data.float()
        w_down = self.lora_layer.down.weight.data.float()
        fused = self.weight.data.float() + lora_scale * (w_up @ w_down)
        self.weight = nn.Parameter(fused.to(self.weight.dtype))  # un-ties from the UNet on purpose
        self.lora_layer = None

    def forward(self, x, scale: float = 1.0):
        out = super().forward(x)
        if self.lora_layer is not None:
            out = out + scale * self.lora_layer(x)
        return out


class LoRACompatibleConv(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_layer: Optional[LoRAConv2dLayer] = None

    def set_lora_layer(self, lora_layer):
        self.lora_layer = lora_layer

    def _fuse_lora(self, lora_scale: float = 1.0):
        if self.lora_layer is None:
            return
        w_up = self.lora_layer.up.weight.data.float().flatten(1)
        w_down = self.lora_layer.down.weight.data.float().flatten(1)
        fusion = (w_up @ w_down).reshape(self.weight.shape)
        self.weight = nn.Parameter((self.weight.data.float() + lora_scale * fusion).to(self.weight.dtype))
        self.lora_layer = None

    def forward(self, x, scale: float = 1.0):
        out = super().forward(x)
        if self.lora_layer is not None:
            out = out + scale * self.lora_layer(x)
        return out


Linear = LoRACompatibleLinear
Conv2d = LoRACompatibleConv


# --------------------------------------------------------------------------------------------
# Time embedding (A.1)
# --------------------------------------------------------------------------------------------
def timestep_sinusoid(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """flip_sin_to_cos=True, freq_shift=0, max_period=10000 -> cat([cos, sin]); always fp32."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    args = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = Linear(in_dim, dim)
        self.linear_2 = Linear(dim, dim)

    def forward(self, t_emb):
        return self.linear_2(F.silu(self.linear_1(t_emb)))


# --------------------------------------------------------------------------------------------
# ResnetBlock2D (A.2)
# --------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_dim: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = Conv2d(cin, cout, 3, 1, 1)
        self.time_emb_proj = Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = Conv2d(cin, cout, 1, 1, 0) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


# --------------------------------------------------------------------------------------------
# Transformer2DModel (A.3)
# --------------------------------------------------------------------------------------------
class Attention(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: Optional[int] = None):
        super().__init__()
        self.heads = heads
        ctx_dim = dim if ctx_dim is None else ctx_dim
        self.to_q = Linear(dim, dim, bias=False)
        self.to_k = Linear(ctx_dim, dim, bias=False)
        self.to_v = Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([Linear(dim, dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, n, c = x.shape
        h = self.heads
        q = self.to_q(x).view(b, n, h, c // h).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], h, c // h).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], h, c // h).transpose(1, 2)
        # softmax(q k^T / sqrt(d)) v, no mask, no dropout (AttnProcessor2_0 -> SDPA)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = Linear(dim, inner * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)  # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6, affine=True)
        self.proj_in = Conv2d(dim, dim, 1, 1, 0)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim)])
        self.proj_out = Conv2d(dim, dim, 1, 1, 0)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        r = x
        x = self.proj_in(self.norm(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            x = blk(x, ctx)
        x = x.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(x) + r


# --------------------------------------------------------------------------------------------
# Down / mid / up blocks (A.4)
# --------------------------------------------------------------------------------------------
class Downsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = Conv2d(c, c, 3, 2, 1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = Conv2d(c, c, 3, 1, 1)

    def forward(self, x, size=None):
        if size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=size, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    """CrossAttnDownBlock2D (has_attn) or DownBlock2D."""

    def __init__(self, cfg: SD15Config, cin: int, cout: int, has_attn: bool, add_down: bool):
        super().__init__()
        self.has_cross_attention = has_attn
        self.resnets = nn.ModuleList(
            [
                ResnetBlock2D(cin if i == 0 else cout, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
                for i in range(cfg.layers_per_block)
            ]
        )
        if has_attn:
            self.attentions = nn.ModuleList(
                [
                    Transformer2DModel(cout, cfg.num_heads, cfg.cross_attention_dim, cfg.norm_num_groups)
                    for _ in range(cfg.layers_per_block)
                ]
            )
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.has_cross_attention:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: SD15Config, c: int):
        super().__init__()
        self.has_cross_attention = True
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(c, c, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)]
        )
        self.attentions = nn.ModuleList(
            [Transformer2DModel(c, cfg.num_heads, cfg.cross_attention_dim, cfg.norm_num_groups)]
        )

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """UpBlock2D / CrossAttnUpBlock2D: 3 resnets over cat[x, skip]."""

    def __init__(self, cfg: SD15Config, cin: int, cout: int, cprev: int, has_attn: bool, add_up: bool):
        super().__init__()
        self.has_cross_attention = has_attn
        n = cfg.layers_per_block + 1
        resnets = []
        for i in range(n):
            res_skip = cin if i == n - 1 else cout
            res_in = cprev if i == 0 else cout
            resnets.append(
                ResnetBlock2D(res_in + res_skip, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
            )
        self.resnets = nn.ModuleList(resnets)
        if has_attn:
            self.attentions = nn.ModuleList(
                [
                    Transformer2DModel(cout, cfg.num_heads, cfg.cross_attention_dim, cfg.norm_num_groups)
                    for _ in range(n)
                ]
            )
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips: List[torch.Tensor], temb, ctx, upsample_size=None):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.has_cross_attention:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x, upsample_size)
        return x


def _expand_timestep(timestep, batch: int, device) -> torch.Tensor:
    """controllora.py:133-148: python number / 0-d tensor -> [B] tensor."""
    if not torch.is_tensor(timestep):
        dtype = torch.float64 if isinstance(timestep, float) else torch.int64
        timestep = torch.tensor([timestep], dtype=dtype, device=device)
    elif timestep.dim() == 0:
        timestep = timestep[None].to(device)
    return timestep.expand(batch)


class _Encoder(nn.Module):
    """conv_in + time embedding + down blocks + mid block: shared by UNet and ControlNet."""

    def _build_encoder(self, cfg: SD15Config):
        boc = cfg.block_out_channels
        self.conv_in = Conv2d(cfg.in_channels, boc[0], 3, 1, 1)
        self.time_embedding = TimestepEmbedding(boc[0], cfg.time_embed_dim)
        downs = []
        cout = boc[0]
        for i, c in enumerate(boc):
            cin, cout = cout, c
            downs.append(DownBlock(cfg, cin, cout, cfg.down_has_attn[i], add_down=i != len(boc) - 1))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = MidBlock(cfg, boc[-1])

    def _time(self, sample, timestep):
        t = _expand_timestep(timestep, sample.shape[0], sample.device)
        t_emb = timestep_sinusoid(t, self.cfg.block_out_channels[0]).to(sample.dtype)
        return self.time_embedding(t_emb)


class UNet2DConditionModel(_Encoder):
    def __init__(self, cfg: Optional[SD15Config] = None):
        super().__init__()
        self.cfg = cfg = cfg or SD15Config()
        boc = cfg.block_out_channels
        self._build_encoder(cfg)
        ups = []
        rev = list(reversed(boc))
        rev_attn = list(reversed(cfg.down_has_attn))
        cout = rev[0]
        for i in range(len(boc)):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            ups.append(UpBlock(cfg, cin, cout, cprev, rev_attn[i], add_up=i != len(boc) - 1))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = Conv2d(boc[0], cfg.out_channels, 3, 1, 1)

    def forward(
        self,
        sample,
        timestep,
        encoder_hidden_states,
        down_block_additional_residuals: Optional[Sequence[torch.Tensor]] = None,
        mid_block_additional_residual: Optional[torch.Tensor] = None,
    ):
        emb = self._time(sample, timestep)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        if down_block_additional_residuals is not None:
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
        x = self.mid_block(x, emb, encoder_hidden_states)
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual
        for i, blk in enumerate(self.up_blocks):
            n = len(blk.resnets)
            blk_skips = skips[-n:]
            skips = skips[:-n]
            up_size = skips[-1].shape[2:] if (blk.upsamplers is not None and skips) else None
            x = blk(x, blk_skips, emb, encoder_hidden_states, up_size)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return x


class ControlNetConditioningEmbedding(nn.Module):
    """openpose-style raw-image embedder (A.4): 3x512x512 -> 320x64x64; conv_out zero-init."""

    def __init__(self, out_ch: int, cond_ch: int, block_out: Sequence[int]):
        super().__init__()
        self.conv_in = nn.Conv2d(cond_ch, block_out[0], 3, padding=1)
        blocks = []
        for i in range(len(block_out) - 1):
            blocks.append(nn.Conv2d(block_out[i], block_out[i], 3, padding=1))
            blocks.append(nn.Conv2d(block_out[i], block_out[i + 1], 3, padding=1, stride=2))
        self.blocks = nn.ModuleList(blocks)
        self.conv_out = nn.Conv2d(block_out[-1], out_ch, 3, padding=1)
        nn.init.zeros_(self.conv_out.weight)
        nn.init.zeros_(self.conv_out.bias)

    def forward(self, c):
        e = F.silu(self.conv_in(c))
        for b in self.blocks:
            e = F.silu(b(e))
        return self.conv_out(e)


class ControlNetModel(_Encoder):
    """diffusers ControlNetModel with CachedControlNetModel.forward semantics
    (/root/reference/model/controllora.py:59-287): the conditioning embedder is skipped when
    `controlnet_cond` already has the latent's spatial size (:199-201)."""

    def __init__(self, cfg: Optional[SD15Config] = None):
        super().__init__()
        self.cfg = cfg = cfg or SD15Config()
        boc = cfg.block_out_channels
        self._build_encoder(cfg)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(
            boc[0], cfg.conditioning_channels, cfg.conditioning_embedding_out_channels
        )
        zc = [boc[0]]
        for i, c in enumerate(boc):
            zc += [c] * cfg.layers_per_block
            if i != len(boc) - 1:
                zc.append(c)
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(c, c, 1) for c in zc])
        self.controlnet_mid_block = nn.Conv2d(boc[-1], boc[-1], 1)
        for m in list(self.controlnet_down_blocks) + [self.controlnet_mid_block]:
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)

    def preprocess_image(self, image):  # controllora.py:289-290
        return self.controlnet_cond_embedding(image)

    def forward(
        self,
        sample,
        timestep,
        encoder_hidden_states,
        controlnet_cond,
        conditioning_scale: float = 1.0,
        guess_mode: bool = False,
        return_dict: bool = False,
    ):
        emb = self._time(sample, timestep)
        x = self.conv_in(sample)
        if controlnet_cond.shape[2:] != x.shape[2:]:
            controlnet_cond = self.controlnet_cond_embedding(controlnet_cond)
        x = x + controlnet_cond
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states)
        down = [zc(s) for s, zc in zip(skips, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(x)
        if guess_mode:  # controllora.py:257-265
            scales = torch.logspace(-1, 0, len(down) + 1, device=sample.device) * conditioning_scale
            down = [d * s for d, s in zip(down, scales)]
            mid = mid * scales[-1]
        else:
            down = [d * conditioning_scale for d in down]
            mid = mid * conditioning_scale
        return down, mid


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
