"""Host mirror of /root/reference/model/edgestyle_pipeline.py: EdgeStyleStableDiffusionControlNetPipeline.

`__call__` keeps the reference's keyword surface (:92-120).  The hot-path subset is implemented
(`prompt_embeds` / `negative_prompt_embeds` / `latents` / cached conditioning embeddings / `guidance_scale`
/ `num_inference_steps` / `num_images_per_prompt` / `eta` / `controlnet_conditioning_scale` /
`control_guidance_start|end` / `generator` (one or a list) / `guess_mode` / `output_type` / `callback_on_step_end`);
`timesteps=` raises the reference's own ValueError (neither of its schedulers takes a custom schedule); the per-call stages either side of the loop run on
`edgestyle_b200.vae.AutoencoderKL` (raw control images of the ControlLoRA nets -> VAE conditioning embedding; final
latents -> image for `output_type` "pt" / "np" / "pil"); CLIP text encoding and the safety checker are outside the
path and raise NotImplementedError instead of being silently ignored.

Per step (reference loop body :434-543): CFG duplicate -> 6 ControlNets + merge -> UNet -> CFG combine ->
scheduler.step.  Here: one captured CUDA graph (DenoiseEngine.step) + one es_cfg_ddim launch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Union

import torch

from ..schedulers import DDIMScheduler
from .controllora import UNet2DConditionModel
from .edgestyle_multicontrolnet import EdgeStyleMultiControlNetModel


@dataclass
class StableDiffusionPipelineOutput:
    images: Any
    nsfw_content_detected: Optional[List[bool]]


class EdgeStyleStableDiffusionControlNetPipeline:
    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet: UNet2DConditionModel = None,
                 controlnet: EdgeStyleMultiControlNetModel = None, scheduler=None, safety_checker=None,
                 feature_extractor=None, image_encoder=None, requires_safety_checker: bool = True,
                 use_graph: bool = True):
        if unet is None or controlnet is None:
            raise ValueError("unet and controlnet are required")
        if controlnet.unet() is not unet:
            raise ValueError("the ControlLoRA nets must be tied to this pipeline's unet (app.py:95-97)")
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.controlnet = unet, controlnet
        self.scheduler = scheduler or DDIMScheduler()
        self.use_graph = use_graph
        self._guidance_scale = 7.5
        self.h2d_bytes = 0  # bytes copied host->device / device->host by the last __call__ (bench.py e2e accounting)
        self.d2h_bytes = 0

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def do_classifier_free_guidance(self):
        g = self._guidance_scale
        return bool((g > 1).all()) if torch.is_tensor(g) else g > 1

    def _to_dev(self, t: torch.Tensor, dev) -> torch.Tensor:
        if t.device.type != "cuda":
            self.h2d_bytes += t.numel() * t.element_size()
            return t.to(dev, non_blocking=True)
        return t

    @torch.no_grad()
    def __call__(self, prompt=None, image=None, height=None, width=None, num_inference_steps: int = 50,
                 timesteps=None, guidance_scale=7.5, negative_prompt=None, num_images_per_prompt: int = 1,
                 eta: float = 0.0, generator=None, latents=None, prompt_embeds=None, negative_prompt_embeds=None,
                 ip_adapter_image=None, output_type: str = "pil", return_dict: bool = True,
                 cross_attention_kwargs=None, controlnet_conditioning_scale: Union[float, List[float]] = 1.0,
                 guess_mode: bool = False, control_guidance_start: Union[float, List[float]] = 0.0,
                 control_guidance_end: Union[float, List[float]] = 1.0, clip_skip=None,
                 callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], **kwargs):
        # ---- arguments outside the hot path: explicit errors, never silently ignored ----
        if prompt is not None or negative_prompt is not None:
            raise NotImplementedError("text encoding is outside the hot path: pass prompt_embeds / negative_prompt_embeds")
        if prompt_embeds is None:
            raise ValueError("prompt_embeds is required")
        if output_type not in ("latent", "pt", "np", "pil"):
            raise ValueError(f"unknown output_type {output_type!r}")
        if output_type != "latent" and self.vae is None:
            raise ValueError('output_type != "latent" needs the pipeline\'s vae (edgestyle_b200.vae.AutoencoderKL)')
        for name, val in (("ip_adapter_image", ip_adapter_image), ("clip_skip", clip_skip)):
            if val is not None:
                raise NotImplementedError(f"{name} is not implemented")
        if timesteps is not None:
            # retrieve_timesteps (:698-706): neither scheduler the reference uses (DDIM, UniPC of diffusers 0.26.3) takes a
            # custom schedule -- the reference raises this ValueError
            raise ValueError(f"The current scheduler class {self.scheduler.__class__}'s `set_timesteps` does not support custom"
                             f" timestep schedules. Please check whether you are using the correct scheduler.")
        if cross_attention_kwargs and cross_attention_kwargs.get("scale", 1.0) != 1.0:
            raise NotImplementedError('cross_attention_kwargs["scale"] != 1')
        if not isinstance(image, (list, tuple)) or len(image) != 6:
            raise ValueError("`image` must be the list of six conditioning tensors")
        self._guidance_scale = guidance_scale
        cfg_on = self.do_classifier_free_guidance  # guidance_scale <= 1 disables CFG (:319,329,443-447): one row per image
        n_prompts = prompt_embeds.shape[0]
        n_per = int(num_images_per_prompt)
        if n_per < 1:
            raise ValueError("num_images_per_prompt must be >= 1")
        n_img = n_prompts * n_per
        if cfg_on and negative_prompt_embeds is None:
            raise ValueError("negative_prompt_embeds is required when guidance_scale > 1")
        if n_per > 1:  # encode_prompt: repeat(1, n, 1).view(bs * n, seq, -1) -- the copies of a prompt are adjacent
            prompt_embeds = prompt_embeds.repeat_interleave(n_per, dim=0)
            if negative_prompt_embeds is not None:
                negative_prompt_embeds = negative_prompt_embeds.repeat_interleave(n_per, dim=0)
        if isinstance(generator, (list, tuple)) and len(generator) != n_img:  # prepare_latents (:613-617)
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                             f" size of {n_img}. Make sure the batch size matches the length of the generators.")
        self.h2d_bytes = self.d2h_bytes = 0
        dev = torch.device("cuda", torch.cuda.current_device())
        B = 2 * n_img if cfg_on else n_img
        nets = 6
        # align control guidance (edgestyle_pipeline.py:264-283)
        if not isinstance(control_guidance_start, list):
            control_guidance_start = [control_guidance_start] * nets
        if not isinstance(control_guidance_end, list):
            control_guidance_end = [control_guidance_end] * nets
        if isinstance(controlnet_conditioning_scale, (int, float)):
            controlnet_conditioning_scale = [float(controlnet_conditioning_scale)] * nets
        # ---- inputs -> device ----
        if cfg_on:
            pe = torch.cat([self._to_dev(negative_prompt_embeds, dev), self._to_dev(prompt_embeds, dev)])  # :330
        else:
            pe = self._to_dev(prompt_embeds, dev)
        c0 = self.unet.config.block_out_channels[0]
        # latent size from the first conditioning entry (cached embedding [*, c0, h, w] or raw image [*, 3, 8h, 8w]);
        # the step engine is built first so that the per-call embedders below share its weights
        h, w = image[0].shape[-2:] if image[0].shape[1] == c0 else (image[0].shape[-2] // 8, image[0].shape[-1] // 8)
        eng = self.controlnet.engine(B, h, w, use_graph=self.use_graph)
        conds = []
        for net, c in zip(self.controlnet.nets, image):
            c = self._to_dev(c, dev)
            # prepare_image (:645-653): one control image serves the whole batch, one per prompt serves its copies
            if c.shape[0] == 1 and n_img > 1:
                c = c.repeat_interleave(n_img, dim=0)
            elif n_per > 1 and c.shape[0] == n_prompts:
                c = c.repeat_interleave(n_per, dim=0)
            if c.shape[1] != c0:
                # raw control image [n, 3, 8h, 8w]: the per-call precompute of prepare_image (:629-664): openpose nets
                # run their ControlNetConditioningEmbedding, ControlLoRA nets the VAE encoder + conv_vae_out.  The
                # reference embeds AFTER the CFG duplication (:657-662), so the two CFG rows of a ControlLoRA net
                # carry independently sampled VAE latents: one encoder pass, `repeats` draws.
                if getattr(net, "controlnet_conditioning_channel_order", "rgb") == "bgr":
                    c = torch.flip(c, dims=[1])
                c = net.preprocess_image(c, repeats=2 if (cfg_on and c.shape[0] == n_img) else 1)
            if cfg_on and c.shape[0] == n_img:  # CFG duplication of the cached embedding (:657-658)
                c = torch.cat([c] * 2)
            if c.shape[0] != B:
                raise ValueError(f"conditioning embedding has {c.shape[0]} rows, expected {B}")
            conds.append(c)
        if tuple(conds[0].shape[-2:]) != (h, w):
            raise ValueError(f"conditioning embeddings are {tuple(conds[0].shape[-2:])}, expected latent size {(h, w)}")
        eng.set_prompt(pe)
        eng.set_conditioning(conds)
        sch = self.scheduler
        ts = sch.set_timesteps(num_inference_steps)
        if latents is None:
            latents = self._randn((n_img, self.unet.config.in_channels, h, w), generator)
        elif latents.shape[0] != n_img:
            raise ValueError(f"latents has {latents.shape[0]} rows, expected {n_img}")
        latents = self._to_dev(latents, dev).to(torch.float32).clone() * sch.init_noise_sigma
        if cfg_on:
            g = guidance_scale if torch.is_tensor(guidance_scale) else torch.full((n_img,), float(guidance_scale))
            eng.guidance.copy_(self._to_dev(g.to(torch.float32), dev).reshape(-1).expand(n_img))
            self.h2d_bytes += 0 if torch.is_tensor(guidance_scale) else 4 * n_img
        # prepare_extra_step_kwargs (:411): eta reaches scheduler.step only if it takes one -- DDIM does, UniPC ignores it
        eta = float(eta) if not hasattr(sch, "device_step") else 0.0
        n_t = len(ts)
        for i, t in enumerate(ts):
            keeps = [1.0 - float(i / n_t < s or (i + 1) / n_t > e)
                     for s, e in zip(control_guidance_start, control_guidance_end)]  # :418-427
            cond_scale = [c * k for c, k in zip(controlnet_conditioning_scale, keeps)]
            x = sch.scale_model_input(torch.cat([latents] * 2) if cfg_on else latents, t)  # :443-450
            # guess_mode (:453-459, 487-497): logspace-scaled ControlNet outputs; under CFG the ControlNets only act on
            # the conditional rows, the unconditional rows keep the plain UNet skips
            eng.step(x, float(t), cond_scale, guess_mode=guess_mode, zero_uncond=guess_mode and cfg_on)
            if hasattr(sch, "device_step"):     # UniPC: x0-prediction + predictor/corrector linear combinations
                sch.device_step(eng.eps_out, latents, eng.guidance if cfg_on else None)
            elif eta != 0.0:                  # stochastic DDIM (the rare path: a few small torch kernels)
                self._ddim_eta_update(latents, eng.eps_out, eng.guidance if cfg_on else None, sch, int(t), eta, generator)
            elif not cfg_on:                  # DDIM without CFG: x' = (a'/a) x + (s' - a' s / a) eps, a = sqrt(abar)
                import math

                from .. import ops
                a_t, a_prev = sch.coefficients(int(t))
                al, sg, alp, sgp = math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_prev), math.sqrt(1 - a_prev)
                ops.lincomb(latents, [(alp / al, latents), (sgp - alp * sg / al, eng.eps_out)])
            else:                             # DDIM: fused CFG + update
                a_t, a_prev = sch.coefficients(int(t))
                eng.cfg_ddim_update(latents, a_t, a_prev)
            self.h2d_bytes += 4 + 16  # timestep + 4 scheduler coefficients
            if callback_on_step_end is not None:  # :523-533
                avail = {"latents": latents, "prompt_embeds": pe[n_img:] if cfg_on else pe,
                         "negative_prompt_embeds": pe[:n_img] if cfg_on else None}
                callback_kwargs = {}
                for k in callback_on_step_end_tensor_inputs:
                    if k not in avail:
                        raise KeyError(f"callback_on_step_end_tensor_inputs: {k!r} is not available (have {sorted(avail)})")
                    callback_kwargs[k] = avail[k]
                out = callback_on_step_end(self, i, t, callback_kwargs) or {}
                latents = out.pop("latents", latents)
                if "prompt_embeds" in out or "negative_prompt_embeds" in out:
                    p_new = out.pop("prompt_embeds", avail["prompt_embeds"])
                    n_new = out.pop("negative_prompt_embeds", avail["negative_prompt_embeds"])
                    pe = torch.cat([self._to_dev(n_new, dev), self._to_dev(p_new, dev)]) if cfg_on else self._to_dev(p_new, dev)
                    eng.set_prompt(pe)
        images = latents
        if output_type != "latent":  # :552-572 (no safety checker: has_nsfw_concept is None, every image denormalised)
            images = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False, generator=generator)[0]
            images = self._postprocess(images, output_type)
        if not return_dict:
            return (images, None)
        return StableDiffusionPipelineOutput(images=images, nsfw_content_detected=None)

    @staticmethod
    def _randn(shape, generator, device="cpu"):
        """diffusers randn_tensor: a list of generators draws one image each; a CPU generator draws on the CPU (the
        result is copied to the device by the caller)."""
        if isinstance(generator, (list, tuple)):
            one = (1,) + tuple(shape[1:])
            return torch.cat([torch.randn(one, generator=g, device=g.device) for g in generator])
        return torch.randn(shape, generator=generator, device=generator.device if generator is not None else device)

    def _ddim_eta_update(self, latents, eps, guidance, sch, t: int, eta: float, generator):
        """DDIMScheduler.step with eta > 0 (formulas (12), (16) of the DDIM paper as diffusers 0.26.3 writes them):
        sigma_t = eta sqrt((1 - a') / (1 - a)) sqrt(1 - a / a'), direction sqrt(1 - a' - sigma_t^2) eps, plus sigma_t z."""
        import math

        a_t, a_prev = sch.coefficients(t)
        if guidance is not None:
            u, c = eps.chunk(2)
            eps = u + guidance.view(-1, 1, 1, 1) * (c - u)
        std = eta * math.sqrt((1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev))
        x0 = (latents - math.sqrt(1 - a_t) * eps) / math.sqrt(a_t)
        z = self._randn(tuple(latents.shape), generator, device=latents.device)
        z = self._to_dev(z, latents.device).to(latents.dtype)
        latents.copy_(math.sqrt(a_prev) * x0 + math.sqrt(max(1 - a_prev - std * std, 0.0)) * eps + std * z)

    def _postprocess(self, image: torch.Tensor, output_type: str):
        """VaeImageProcessor.postprocess (diffusers image_processor.py): denormalise to [0, 1], then "pt" (NCHW
        tensor), "np" (NHWC float32 array, a device->host copy) or "pil".  Host-side image formatting of the decoded
        tensor, outside the kernels."""
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return image
        arr = image.cpu().permute(0, 2, 3, 1).float().numpy()
        self.d2h_bytes += arr.nbytes
        if output_type == "np":
            return arr
        from PIL import Image

        return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
