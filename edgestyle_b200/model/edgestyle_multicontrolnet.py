"""Host mirror of /root/reference/model/edgestyle_multicontrolnet.py: EdgeStyleMultiControlNetModel.

Same constructor shape (`controlnets` list -> `.nets`), same `forward` signature and return convention
(:116-171: always a `(down_block_res_samples, mid_block_res_sample)` tuple, `return_dict` ignored), merge-block
state dict under `multi_controlnet_down_blocks.* / multi_controlnet_mid_block.*` (:173-193).  The six
sequential ControlNet calls + interleave + 13 ControlNetBlocks of the reference become two batched encoder
passes + a fused merge kernel in `edgestyle_b200.engine.DenoiseEngine`.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import torch

from .. import config as C
from ..engine import SUPPORTED_PATTERN, DenoiseEngine
from .controllora import CachedControlNetModel, ControlLoRAModel, UNet2DConditionModel, _check_spec


class EdgeStyleMultiControlNetModel:
    def __init__(self, controlnets: Sequence[CachedControlNetModel], merge_state_dict: Optional[Mapping] = None,
                 latent_hw: Tuple[int, int] = (64, 64), dtype=torch.float16, n_text: int = 77):
        self.nets: List[CachedControlNetModel] = list(controlnets)
        if len(self.nets) != 6:
            raise NotImplementedError("EdgeStyle uses exactly six control branches (app.py:86-94)")
        a, p0, c0, p1, c1, p2 = self.nets
        ok = (isinstance(a, ControlLoRAModel) and isinstance(c0, ControlLoRAModel) and c0 is c1 and a is not c0
              and p0 is p1 is p2 and not isinstance(p0, ControlLoRAModel))
        if not ok:
            raise NotImplementedError(
                f"only the reference's net pattern {SUPPORTED_PATTERN} = [loraA, pose, loraB, pose, loraB, pose] with "
                "shared module objects (app.py:40,86-94) is supported")
        self.config = a.config
        self.latent_hw = tuple(latent_hw)
        self.dtype, self.n_text = dtype, n_text
        spec = C.merge_spec(self.config, *self.latent_hw)
        if merge_state_dict is None:  # nn.Conv2d / nn.LayerNorm default init, as the reference's constructor
            g = torch.Generator().manual_seed(0)
            merge_state_dict = OrderedDict()
            for k, shape in spec.items():
                if "normalization" in k:
                    merge_state_dict[k] = torch.ones(shape) if k.endswith("weight") else torch.zeros(shape)
                else:
                    fan_in = 2 if "first_conv" in k else (3 if "second_conv" in k else 1)
                    merge_state_dict[k] = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        _check_spec(merge_state_dict, spec, "EdgeStyleMultiControlNetModel")
        self._merge_sd = OrderedDict(merge_state_dict)
        for slot, net in enumerate(self.nets):
            net._owner = self
            if slot not in net._slots:
                net._slots.append(slot)
        self._engines: Dict[tuple, DenoiseEngine] = {}
        self._fused = False  # fuse(): engines built afterwards use fused LoRA weight copies regardless of ES_LORA

    # ---------------------------------------------------------------------------------------
    def state_dict(self):
        return self._merge_sd

    def load_state_dict(self, state_dict: Mapping, strict: bool = True):
        """Merge-block tensors (edgestyle_multicontrolnet.py:173-193); cached engines are rebuilt on the next call."""
        spec = C.merge_spec(self.config, *self.latent_hw)
        if strict:
            _check_spec(state_dict, spec, "EdgeStyleMultiControlNetModel")
        for k, v in state_dict.items():
            if k in spec:
                self._merge_sd[k] = v
        self._invalidate_engines()

    def _invalidate_engines(self):
        """Engines hold packed device copies of the nets' weights and captured CUDA graphs: a weight reload / re-tie of
        any net (ControlLoRAModel.load_state_dict / tie_weights) must drop them (the reference's modules are live)."""
        self._engines.clear()

    def lora_group(self, net) -> Optional[int]:
        """LoRA weight group of a registered net inside the engine: 0 = agnostic set, 1 = clothes set, None = plain."""
        if not net.uses_lora:
            return None
        return 0 if net is self.nets[0] else 1

    def unet(self) -> UNet2DConditionModel:
        u = self.nets[0]._unet
        if u is None or self.nets[2]._unet is not u:
            raise RuntimeError("call net.tie_weights(unet) on both ControlLoRA nets first (app.py:95-97)")
        return u

    def engine(self, rows: int, h: int, w: int, use_graph: bool = False) -> DenoiseEngine:
        if (h, w) != self.latent_hw:
            raise ValueError(f"merge blocks were built for latent {self.latent_hw}, got {(h, w)} "
                             "(LayerNorm([C,H,W]) ties the model to one resolution, SURVEY.md F7)")
        key = (rows, h, w, use_graph)
        eng = self._engines.get(key)
        if eng is None:
            eng = DenoiseEngine(self.config, self.unet().state_dict(),
                                [self.nets[0].state_dict(), self.nets[2].state_dict()], self.nets[1].state_dict(),
                                self._merge_sd, rows=rows, h=h, w=w, dtype=self.dtype, n_text=self.n_text,
                                use_graph=use_graph, fuse_lora=True if self._fused else None)
            self._engines[key] = eng
        return eng

    def embed_engine(self, rows: int, h: int, w: int) -> DenoiseEngine:
        """Engine for the per-call conditioning embedders (`embed_openpose`, `embed_vae_latent`): they only use the
        packed embedder / conv_in weights and scratch keyed by their own batch size, so any engine already built for
        this latent size serves (no second weight replica for a different row count)."""
        for (_, hh, ww, _), eng in self._engines.items():
            if (hh, ww) == (h, w):
                return eng
        return self.engine(rows, h, w)

    # -- reference surface ----------------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond: List[torch.Tensor],
                conditioning_scale: List[float], class_labels=None, timestep_cond=None, attention_mask=None,
                added_cond_kwargs=None, cross_attention_kwargs=None, guess_mode: bool = False,
                return_dict: bool = True):
        for name, val in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                          ("attention_mask", attention_mask), ("added_cond_kwargs", added_cond_kwargs)):
            if val is not None:
                raise NotImplementedError(f"{name} is not used by SD1.5 and not implemented")
        if len(controlnet_cond) != 6 or len(conditioning_scale) != 6:
            raise ValueError("expected six conditioning tensors and six scales")
        B, _, h, w = sample.shape
        eng = self.engine(B, h, w)
        eng.set_prompt(encoder_hidden_states)
        eng.set_conditioning(self.embed_conditioning(controlnet_cond, (h, w)))
        # guess_mode: each net scales its outputs by logspace(-1, 0, 13) * scale (controllora.py:257-265) before the merge
        down, mid = eng.residuals(sample, timestep, conditioning_scale, guess_mode=guess_mode)
        return down, mid

    __call__ = forward

    def embed_conditioning(self, controlnet_cond: List[torch.Tensor], latent_hw) -> List[torch.Tensor]:
        """Raw control images (spatial size != latent) go through their net's conditioning embedder
        (controllora.py:199-201 / preprocess_image :289-290); latent-sized entries are the cached embeddings."""
        out = []
        for net, c in zip(self.nets, controlnet_cond):
            if tuple(c.shape[2:]) != tuple(latent_hw):
                if getattr(net, "controlnet_conditioning_channel_order", "rgb") == "bgr":
                    c = torch.flip(c, dims=[1])
                c = net.preprocess_image(c)
            out.append(c)
        return out

    def _single_forward(self, net, sample, timestep, ehs, cond, scale, guess_mode):
        B, _, h, w = sample.shape
        eng = self.engine(B, h, w)
        return eng.single_controlnet(self.lora_group(net), sample, timestep, ehs, cond, scale, guess_mode)

    def fuse(self):  # noqa: D401
        """edgestyle_multicontrolnet.py:284-287 replaces every ControlLoRA net by `net.fuse()`.  Here the nets stay
        (so their conditioning embeddings remain cacheable, which the reference loses after fusing: SURVEY.md appendix,
        quirk 11) and the engine is pinned to one fused weight copy `W + up @ down` per LoRA group -- the same
        arithmetic as running the fused nets, whatever ES_LORA says."""
        if not self._fused:
            self._fused = True
            self._engines.clear()

    # -- checkpoint directory format of the reference (edgestyle_multicontrolnet.py:213-282, 289-430) ----------------
    def save_pretrained(self, save_directory, save_pattern: Optional[Sequence[Optional[int]]] = None, **_):
        """<dir>/diffusion_pytorch_model.safetensors = merge blocks only (:173-193);
        <dir>/controlnet_{idx}/ for every distinct idx of `save_pattern` (:242-281)."""
        import os

        from safetensors.torch import save_file

        from .controllora import WEIGHTS_NAME

        if os.path.isfile(save_directory):
            raise ValueError(f"Provided path ({save_directory}) should be a directory, not a file")
        os.makedirs(save_directory, exist_ok=True)
        save_file({k: v.detach().cpu().contiguous() for k, v in self._merge_sd.items()},
                  os.path.join(save_directory, WEIGHTS_NAME), metadata={"format": "pt"})
        save_pattern = list(save_pattern) if save_pattern is not None else [None] * len(self.nets)
        saved = []
        for i, net in enumerate(self.nets):
            if save_pattern[i] is not None and save_pattern[i] not in saved:
                net.save_pretrained(os.path.join(save_directory, f"controlnet_{save_pattern[i]}"))
                saved.append(save_pattern[i])

    @classmethod
    def from_pretrained(cls, pretrained_model_path, *, load_pattern=None, static_controlnets=None,
                        controlnet_class=ControlLoRAModel, vae=None, torch_dtype=None, latent_hw=(64, 64),
                        variant: Optional[str] = None, **kwargs):
        import os

        from safetensors.torch import load_file

        from .controllora import WEIGHTS_NAME

        if not os.path.isdir(pretrained_model_path):
            raise ValueError(f"Provided path ({pretrained_model_path}) should be a directory")
        if variant is not None:
            raise NotImplementedError("weight-file variants (e.g. 'fp16') are not implemented: the mirrors read "
                                      "diffusion_pytorch_model.safetensors")
        if load_pattern is None:
            raise ValueError("load_pattern must be provided")
        static_controlnets = static_controlnets or [None] * len(load_pattern)
        loaded, nets = {}, []
        for i, load in enumerate(load_pattern):
            if load is not None:
                if load not in loaded:
                    loaded[load] = controlnet_class.from_pretrained(os.path.join(pretrained_model_path, f"controlnet_{load}"))
                    if vae is not None and hasattr(loaded[load], "set_autoencoder"):
                        loaded[load].set_autoencoder(vae)
                nets.append(loaded[load])
            else:
                nets.append(static_controlnets[i])
        for i, n in enumerate(nets):
            if n is None:
                raise ValueError(f"All controlnets must be provided. controlnet {i} is None.")
        merge_sd = load_file(os.path.join(pretrained_model_path, WEIGHTS_NAME))
        if torch_dtype is not None and not isinstance(torch_dtype, torch.dtype):
            raise ValueError(f"{torch_dtype} needs to be of type `torch.dtype`")
        return cls(nets, merge_sd, latent_hw=latent_hw, dtype=torch_dtype or torch.float16)
