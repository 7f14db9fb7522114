"""Fused denoise step + synthetic model/input construction (oracle; test infrastructure).

* `fused_step` mirrors OnnxUNetAndControlnets.forward (/root/reference/export_onnx.py:43-74).
* `denoise` mirrors the loop body of EdgeStyleStableDiffusionControlNetPipeline.__call__
  (/root/reference/model/edgestyle_pipeline.py:434-543): CFG duplicate, 6 ControlNets, UNet with
  residuals, CFG combine (:513-517), scheduler.step (:520-522).
* `build_models` / `synthetic_inputs` implement the seeded synthetic workload of SURVEY.md 8(d):
  net pattern [agn, pose, clo, pose, clo, pose] (/root/reference/app.py:40,86-94).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from .controllora import ControlLoRAModel
from .merge import EdgeStyleMultiControlNetModel
from .schedulers import DDIMScheduler
from .sd15 import ControlNetModel, SD15Config, UNet2DConditionModel


@dataclass
class Models:
    cfg: SD15Config
    unet: UNet2DConditionModel
    lora_agnostic: ControlLoRAModel
    lora_clothes: ControlLoRAModel
    openpose: ControlNetModel
    controlnet: EdgeStyleMultiControlNetModel  # nets = [agn, pose, clo, pose, clo, pose]


def _rerandomise_zero_inits(net: nn.Module, gen: torch.Generator, std: float = 0.02):
    """Zero-initialised tensors (zero-convs, LoRA up, cond-embed conv_out) would make all residuals 0."""
    for name, p in net.named_parameters():
        if (name.startswith("controlnet_down_blocks") or name.startswith("controlnet_mid_block")
                or name.endswith("lora_layer.up.weight") or name.startswith("controlnet_cond_embedding.conv_out")):
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=gen) * std)


def _perturb_norm_affine(net: nn.Module, gen: torch.Generator, std: float = 0.05):
    for m in net.modules():
        if isinstance(m, (nn.GroupNorm, nn.LayerNorm)) and m.weight is not None:
            with torch.no_grad():
                m.weight.add_(torch.randn(m.weight.shape, generator=gen) * std)
                m.bias.add_(torch.randn(m.bias.shape, generator=gen) * std)


def build_models(cfg: Optional[SD15Config] = None, latent_hw: Tuple[int, int] = (64, 64), rank: int = 32,
                 seed: int = 0, lora_conv2d_rank: int = 0) -> Models:
    cfg = cfg or SD15Config()
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(cfg)
    agn = ControlLoRAModel(cfg, lora_linear_rank=rank, lora_conv2d_rank=lora_conv2d_rank)
    clo = ControlLoRAModel(cfg, lora_linear_rank=rank, lora_conv2d_rank=lora_conv2d_rank)
    torch.manual_seed(seed + 2)
    pose = ControlNetModel(cfg)
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], cfg, latent_hw)
    g = torch.Generator().manual_seed(seed + 1)
    # perturb norms BEFORE tying so that tied nets see the UNet's perturbed parameters
    _perturb_norm_affine(unet, g)
    _perturb_norm_affine(pose, g)
    for blk in list(multi.multi_controlnet_down_blocks) + [multi.multi_controlnet_mid_block]:
        _perturb_norm_affine(blk, g)
    agn.tie_weights(unet)
    clo.tie_weights(unet)
    for net in (agn, clo, pose):
        _rerandomise_zero_inits(net, g)
    for m in (unet, multi):
        m.eval()
        for p in m.parameters():
            p.requires_grad_(False)
    return Models(cfg, unet, agn, clo, pose, multi)


@dataclass
class StepInputs:
    latents: torch.Tensor  # [B/2, 4, h, w]
    prompt_embeds: torch.Tensor  # [B, 77, ctx] (negative rows first, edgestyle_pipeline.py:330)
    conds: List[torch.Tensor]  # 6 x [B, C0, h, w] cached cond embeddings (already CFG-duplicated)
    conditioning_scale: List[float]


def synthetic_inputs(cfg: SD15Config, images: int = 1, h: int = 64, w: int = 64, seed: int = 1234,
                     n_text: int = 77) -> StepInputs:
    g = torch.Generator().manual_seed(seed)
    B = 2 * images
    latents = torch.randn(images, cfg.in_channels, h, w, generator=g)
    pe = torch.randn(B, n_text, cfg.cross_attention_dim, generator=g)
    conds = [torch.randn(B, cfg.block_out_channels[0], h, w, generator=g) * 0.5 for _ in range(6)]
    return StepInputs(latents, pe, conds, [1.0] * 6)


@torch.no_grad()
def fused_step(models: Models, sample, timestep, encoder_hidden_states, conditioning_scale: Sequence[float],
               conds: Sequence[torch.Tensor]) -> torch.Tensor:
    down, mid = models.controlnet(sample, timestep, encoder_hidden_states, list(conds), list(conditioning_scale),
                                  return_dict=False)
    return models.unet(sample, timestep, encoder_hidden_states,
                       down_block_additional_residuals=down, mid_block_additional_residual=mid)


def cfg_combine(noise_pred: torch.Tensor, guidance_scale) -> torch.Tensor:
    """edgestyle_pipeline.py:513-517.  `guidance_scale` may be a scalar or a per-image vector."""
    u, c = noise_pred.chunk(2)
    if torch.is_tensor(guidance_scale):
        guidance_scale = guidance_scale.view(-1, 1, 1, 1).to(u.dtype)
    return u + guidance_scale * (c - u)


@torch.no_grad()
def denoise(models: Models, inp: StepInputs, num_inference_steps: int = 20, guidance_scale=4.5,
            scheduler=None, return_eps: bool = False, override_latents: Optional[Sequence[torch.Tensor]] = None):
    """Free-running (or, with `override_latents`, teacher-forced) denoise loop.  Returns final latents
    (and per-step (latent_in, eps) if `return_eps`)."""
    sched = scheduler or DDIMScheduler()
    ts = sched.set_timesteps(num_inference_steps)
    latents = inp.latents * sched.init_noise_sigma
    trace = []
    for i, t in enumerate(ts):
        if override_latents is not None:
            latents = override_latents[i]
        x = sched.scale_model_input(torch.cat([latents] * 2), t)
        eps = fused_step(models, x, t, inp.prompt_embeds, inp.conditioning_scale, inp.conds)
        e = cfg_combine(eps, guidance_scale)
        if return_eps:
            trace.append((latents.clone(), eps.clone()))
        latents = sched.step(e, t, latents)
    return (latents, trace) if return_eps else latents


def to_(models: Models, device=None, dtype=None) -> Models:
    models.unet.to(device=device, dtype=dtype)
    models.controlnet.to(device=device, dtype=dtype)
    # re-tie: Module.to() keeps Parameter identity for in-place moves, nothing else to do
    return models
