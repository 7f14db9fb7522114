"""Synthetic random-init weights in diffusers state-dict layout (no network, no checkpoints): used by
bench.py and smoke tests.  Distributions follow PyTorch's default inits (kaiming-uniform, bound
1/sqrt(fan_in)); zero-initialised tensors of the reference (zero-convs, LoRA up) get N(0, 0.02^2) so the
ControlNet branches contribute."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import config as C


def _fill(spec: Dict[str, Tuple[int, ...]], gen: torch.Generator, device) -> Dict[str, torch.Tensor]:
    out = {}
    for name, shape in spec.items():
        is_norm = (".norm" in name or name.startswith("conv_norm_out") or "normalization" in name)
        if name.endswith("lora_layer.up.weight") or name.startswith("controlnet_down_blocks") \
                or name.startswith("controlnet_mid_block") or name.startswith("controlnet_cond_embedding.conv_out"):
            t = torch.randn(shape, generator=gen, device=device) * 0.02
        elif name.endswith("lora_layer.down.weight"):
            t = torch.randn(shape, generator=gen, device=device) / shape[0]
        elif is_norm:
            base = 1.0 if name.endswith(".weight") else 0.0
            t = base + torch.randn(shape, generator=gen, device=device) * 0.05
        else:
            if name.endswith(".weight"):
                fan_in = math.prod(shape[1:]) if len(shape) > 1 else shape[0]
            else:
                wshape = spec.get(name[: -len(".bias")] + ".weight", shape)
                fan_in = math.prod(wshape[1:]) if len(wshape) > 1 else wshape[0]
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            t = (torch.rand(shape, generator=gen, device=device) * 2 - 1) * bound
        out[name] = t
    return out


def synth_state_dicts(cfg: C.UNetConfig, h: int, w: int, rank: int = 32, seed: int = 0, device="cpu",
                      conv_rank: int = 0):
    """Returns dict(unet=..., lora=[agn, clo], pose=..., merge=...) of fp32 tensors."""
    gen = torch.Generator(device=device).manual_seed(seed)
    unet = _fill(C.unet_spec(cfg), gen, device)
    loras = []
    for _ in range(2):
        spec = dict(C.lora_spec(cfg, rank, conv_rank))
        spec.update(C.controlnet_extra_spec(cfg, with_embedder=False))
        loras.append(_fill(spec, gen, device))
    pose_spec = dict(C.encoder_spec(cfg))
    pose_spec.update(C.controlnet_extra_spec(cfg, with_embedder=False))
    pose = _fill(pose_spec, gen, device)
    merge = _fill(C.merge_spec(cfg, h, w), gen, device)
    return {"unet": unet, "lora": loras, "pose": pose, "merge": merge}
