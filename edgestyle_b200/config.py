"""Model hyper-parameters of the path (SD1.5 `unet/config.json`) and the parameter
name/shape specification in diffusers state-dict layout (A.7).  No dependency on the oracle."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    num_heads: int = 8
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)
    conditioning_embedding_out_channels: Tuple[int, ...] = (16, 32, 96, 256)
    conditioning_channels: int = 3

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @classmethod
    def from_any(cls, cfg) -> "UNetConfig":
        """Accept any object with the same attribute names (e.g. the oracle's SD15Config)."""
        if isinstance(cfg, cls):
            return cfg
        names = [f for f in cls.__dataclass_fields__]
        return cls(**{n: tuple(getattr(cfg, n)) if isinstance(getattr(cfg, n), (list, tuple)) else getattr(cfg, n)
                      for n in names})


def level_sizes(h: int, w: int, levels: int) -> List[Tuple[int, int]]:
    out = []
    for _ in range(levels):
        out.append((h, w))
        h, w = (h + 1) // 2, (w + 1) // 2
    return out


def residual_shapes(cfg: UNetConfig, h: int, w: int) -> List[Tuple[int, int, int]]:
    """(C, H, W) of the 12 down residuals + mid (reference table: model/edgestyle_multicontrolnet.py:73-102)."""
    boc = cfg.block_out_channels
    shapes = [(boc[0], h, w)]
    for i, c in enumerate(boc):
        shapes += [(c, h, w)] * cfg.layers_per_block
        if i != len(boc) - 1:
            h, w = (h + 1) // 2, (w + 1) // 2
            shapes.append((c, h, w))
    shapes.append((boc[-1], h, w))
    return shapes


# ------------------------------------------------------------------------------------------------
# Parameter specification (names + shapes), used by the synthetic-weight generator and by tests that
# compare it with the oracle's state_dict().
# ------------------------------------------------------------------------------------------------
def _resnet(p: str, cin: int, cout: int, temb: int) -> Dict[str, Tuple[int, ...]]:
    d = {
        f"{p}.norm1.weight": (cin,), f"{p}.norm1.bias": (cin,),
        f"{p}.conv1.weight": (cout, cin, 3, 3), f"{p}.conv1.bias": (cout,),
        f"{p}.time_emb_proj.weight": (cout, temb), f"{p}.time_emb_proj.bias": (cout,),
        f"{p}.norm2.weight": (cout,), f"{p}.norm2.bias": (cout,),
        f"{p}.conv2.weight": (cout, cout, 3, 3), f"{p}.conv2.bias": (cout,),
    }
    if cin != cout:
        d[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        d[f"{p}.conv_shortcut.bias"] = (cout,)
    return d


def _transformer(p: str, c: int, ctx: int) -> Dict[str, Tuple[int, ...]]:
    t = f"{p}.transformer_blocks.0"
    d = {
        f"{p}.norm.weight": (c,), f"{p}.norm.bias": (c,),
        f"{p}.proj_in.weight": (c, c, 1, 1), f"{p}.proj_in.bias": (c,),
        f"{p}.proj_out.weight": (c, c, 1, 1), f"{p}.proj_out.bias": (c,),
    }
    for n in ("norm1", "norm2", "norm3"):
        d[f"{t}.{n}.weight"] = (c,)
        d[f"{t}.{n}.bias"] = (c,)
    for a, kdim in (("attn1", c), ("attn2", ctx)):
        d[f"{t}.{a}.to_q.weight"] = (c, c)
        d[f"{t}.{a}.to_k.weight"] = (c, kdim)
        d[f"{t}.{a}.to_v.weight"] = (c, kdim)
        d[f"{t}.{a}.to_out.0.weight"] = (c, c)
        d[f"{t}.{a}.to_out.0.bias"] = (c,)
    d[f"{t}.ff.net.0.proj.weight"] = (8 * c, c)
    d[f"{t}.ff.net.0.proj.bias"] = (8 * c,)
    d[f"{t}.ff.net.2.weight"] = (c, 4 * c)
    d[f"{t}.ff.net.2.bias"] = (c,)
    return d


def encoder_spec(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    boc, temb = cfg.block_out_channels, cfg.time_embed_dim
    d = {
        "conv_in.weight": (boc[0], cfg.in_channels, 3, 3), "conv_in.bias": (boc[0],),
        "time_embedding.linear_1.weight": (temb, boc[0]), "time_embedding.linear_1.bias": (temb,),
        "time_embedding.linear_2.weight": (temb, temb), "time_embedding.linear_2.bias": (temb,),
    }
    cout = boc[0]
    for i, c in enumerate(boc):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block):
            d.update(_resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb))
            if cfg.down_has_attn[i]:
                d.update(_transformer(f"down_blocks.{i}.attentions.{j}", cout, cfg.cross_attention_dim))
        if i != len(boc) - 1:
            d[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            d[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
    c = boc[-1]
    d.update(_resnet("mid_block.resnets.0", c, c, temb))
    d.update(_transformer("mid_block.attentions.0", c, cfg.cross_attention_dim))
    d.update(_resnet("mid_block.resnets.1", c, c, temb))
    return d


def unet_spec(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    boc, temb = cfg.block_out_channels, cfg.time_embed_dim
    d = encoder_spec(cfg)
    rev = list(reversed(boc))
    rev_attn = list(reversed(cfg.down_has_attn))
    n = cfg.layers_per_block + 1
    cout = rev[0]
    for i in range(len(boc)):
        cprev, cout = cout, rev[i]
        cin = rev[min(i + 1, len(boc) - 1)]
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = cprev if j == 0 else cout
            d.update(_resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb))
            if rev_attn[i]:
                d.update(_transformer(f"up_blocks.{i}.attentions.{j}", cout, cfg.cross_attention_dim))
        if i != len(boc) - 1:
            d[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            d[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    d["conv_norm_out.weight"] = (boc[0],)
    d["conv_norm_out.bias"] = (boc[0],)
    d["conv_out.weight"] = (cfg.out_channels, boc[0], 3, 3)
    d["conv_out.bias"] = (cfg.out_channels,)
    return d


def zero_conv_channels(cfg: UNetConfig) -> List[int]:
    boc = cfg.block_out_channels
    zc = [boc[0]]
    for i, c in enumerate(boc):
        zc += [c] * cfg.layers_per_block
        if i != len(boc) - 1:
            zc.append(c)
    return zc


def controlnet_extra_spec(cfg: UNetConfig, with_embedder: bool) -> Dict[str, Tuple[int, ...]]:
    d = {}
    for i, c in enumerate(zero_conv_channels(cfg)):
        d[f"controlnet_down_blocks.{i}.weight"] = (c, c, 1, 1)
        d[f"controlnet_down_blocks.{i}.bias"] = (c,)
    c = cfg.block_out_channels[-1]
    d["controlnet_mid_block.weight"] = (c, c, 1, 1)
    d["controlnet_mid_block.bias"] = (c,)
    if with_embedder:
        e = cfg.conditioning_embedding_out_channels
        p = "controlnet_cond_embedding"
        d[f"{p}.conv_in.weight"] = (e[0], cfg.conditioning_channels, 3, 3)
        d[f"{p}.conv_in.bias"] = (e[0],)
        for i in range(len(e) - 1):
            d[f"{p}.blocks.{2 * i}.weight"] = (e[i], e[i], 3, 3)
            d[f"{p}.blocks.{2 * i}.bias"] = (e[i],)
            d[f"{p}.blocks.{2 * i + 1}.weight"] = (e[i + 1], e[i], 3, 3)
            d[f"{p}.blocks.{2 * i + 1}.bias"] = (e[i + 1],)
        d[f"{p}.conv_out.weight"] = (cfg.block_out_channels[0], e[-1], 3, 3)
        d[f"{p}.conv_out.bias"] = (cfg.block_out_channels[0],)
    return d


def lora_linear_names(cfg: UNetConfig) -> List[str]:
    """Every nn.Linear under ControlLoRAModel._skip_layers (/root/reference/model/controllora.py:443-450,529-593)."""
    return [k[: -len(".weight")] for k, s in encoder_spec(cfg).items() if k.endswith(".weight") and len(s) == 2]


def lora_conv_names(cfg: UNetConfig) -> List[str]:
    """Every nn.Conv2d under ControlLoRAModel._skip_layers: conv_in, resnet conv1 / conv2 / conv_shortcut, the
    down-samplers and the transformers' 1x1 proj_in / proj_out (/root/reference/model/controllora.py:538-575)."""
    return [k[: -len(".weight")] for k, s in encoder_spec(cfg).items() if k.endswith(".weight") and len(s) == 4]


def lora_spec(cfg: UNetConfig, rank: int, conv_rank: int = 0) -> Dict[str, Tuple[int, ...]]:
    """LoRA tensors a ControlLoRAModel owns.  conv_rank > 0 adds a LoRAConv2dLayer on every convolution -- whose
    rank is `rank` (lora_linear_rank), not conv_rank: the reference passes rank=lora_linear_rank at controllora.py:569
    and lora_conv2d_rank only switches the branch on."""
    enc = encoder_spec(cfg)
    d = {}
    for n in lora_linear_names(cfg):
        out_f, in_f = enc[n + ".weight"]
        d[f"{n}.lora_layer.down.weight"] = (rank, in_f)
        d[f"{n}.lora_layer.up.weight"] = (out_f, rank)
    if conv_rank > 0:
        for n in lora_conv_names(cfg):
            cout, cin, kh, kw = enc[n + ".weight"]
            d[f"{n}.lora_layer.down.weight"] = (rank, cin, kh, kw)
            d[f"{n}.lora_layer.up.weight"] = (cout, rank, 1, 1)
    return d


def merge_spec(cfg: UNetConfig, h: int, w: int, n_nets: int = 6) -> Dict[str, Tuple[int, ...]]:
    d = {}
    shapes = residual_shapes(cfg, h, w)
    names = [f"multi_controlnet_down_blocks.{i}" for i in range(len(shapes) - 1)] + ["multi_controlnet_mid_block"]
    for p, (c, hh, ww) in zip(names, shapes):
        half = c * n_nets // 2
        d[f"{p}.first_conv.weight"] = (half, 2, 1, 1)
        d[f"{p}.first_conv.bias"] = (half,)
        d[f"{p}.first_normalization.weight"] = (half, hh, ww)
        d[f"{p}.first_normalization.bias"] = (half, hh, ww)
        d[f"{p}.second_conv.weight"] = (c, n_nets // 2, 1, 1)
        d[f"{p}.second_conv.bias"] = (c,)
        d[f"{p}.second_normalization.weight"] = (c, hh, ww)
        d[f"{p}.second_normalization.bias"] = (c, hh, ww)
        d[f"{p}.third_conv.weight"] = (c, 1, 1, 1)
        d[f"{p}.third_conv.bias"] = (c,)
    return d
