"""Host mirror of /root/reference/model/controllora.py for the denoise hot path.

Same class names, `forward` signature (:59-75), `preprocess_image` (:289-290), `from_unet` (:644-725),
`tie_weights` (:623-632), `state_dict` / `load_state_dict` filter semantics (:600-614) and `fuse_lora`
(:728-737).  The classes are weight containers (diffusers-layout state dicts); the arithmetic
runs in the CUDA engine that the owning EdgeStyleMultiControlNetModel builds.  There is no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Dict, List, Mapping, Optional, Tuple, Union

import torch

from .. import config as C

_SKIP_LAYERS = ["conv_in", "time_proj", "time_embedding", "class_embedding", "down_blocks", "mid_block"]


@dataclass
class ControlNetOutput:
    down_block_res_samples: Tuple[torch.Tensor, ...]
    mid_block_res_sample: torch.Tensor


WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"  # diffusers' SAFETENSORS_WEIGHTS_NAME
CONFIG_NAME = "config.json"


def _save_dir(directory, state_dict, config: dict):
    import json
    import os

    from safetensors.torch import save_file

    if os.path.isfile(directory):
        raise ValueError(f"Provided path ({directory}) should be a directory, not a file")
    os.makedirs(directory, exist_ok=True)
    save_file({k: v.detach().cpu().contiguous().clone() for k, v in state_dict.items()}, os.path.join(directory, WEIGHTS_NAME),
              metadata={"format": "pt"})
    json.dump(config, open(os.path.join(directory, CONFIG_NAME), "w"), indent=2)


def _load_dir(directory):
    import json
    import os

    from safetensors.torch import load_file

    if not os.path.isdir(directory):
        raise ValueError(f"Provided path ({directory}) should be a directory")
    cfg_path = os.path.join(directory, CONFIG_NAME)
    config = json.load(open(cfg_path)) if os.path.exists(cfg_path) else {}
    return load_file(os.path.join(directory, WEIGHTS_NAME)), config


def _config_dict(cfg: C.UNetConfig, **extra) -> dict:
    d = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.__dict__.items()}
    d.update(extra)
    return d


def _config_from_dict(d: dict) -> C.UNetConfig:
    names = C.UNetConfig.__dataclass_fields__
    # diffusers' own config.json uses `attention_head_dim: 8` for "8 heads" 
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items() if k in names}
    if "num_heads" not in kw and "attention_head_dim" in d and isinstance(d["attention_head_dim"], int):
        kw["num_heads"] = d["attention_head_dim"]
    if "down_has_attn" not in kw and "down_block_types" in d:  # diffusers' own config.json names the block classes
        known = {"CrossAttnDownBlock2D": True, "DownBlock2D": False}
        bad = [t for t in d["down_block_types"] if t not in known]
        if bad:
            raise NotImplementedError(f"down_block_types {bad}: the engine implements CrossAttnDownBlock2D / DownBlock2D")
        kw["down_has_attn"] = tuple(known[t] for t in d["down_block_types"])
    return C.UNetConfig(**kw)


def _check_spec(sd: Mapping[str, torch.Tensor], spec: Dict[str, Tuple[int, ...]], what: str, allow_extra=()):
    missing = [k for k in spec if k not in sd]
    if missing:
        raise KeyError(f"{what}: missing {len(missing)} keys, e.g. {missing[:3]}")
    for k, shape in spec.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise ValueError(f"{what}: {k} has shape {tuple(sd[k].shape)}, expected {shape}")
    extra = [k for k in sd if k not in spec and not any(k.startswith(p) for p in allow_extra)]
    if extra:
        raise KeyError(f"{what}: unexpected keys, e.g. {extra[:3]}")


class UNet2DConditionModel:
    """Weight container for the SD1.5 UNet (diffusers key names)."""

    def __init__(self, config: C.UNetConfig, state_dict: Mapping[str, torch.Tensor]):
        self.config = C.UNetConfig.from_any(config)
        _check_spec(state_dict, C.unet_spec(self.config), "UNet2DConditionModel")
        self._sd = OrderedDict(state_dict)

    def state_dict(self):
        return self._sd

    def save_pretrained(self, directory):
        _save_dir(directory, self._sd, _config_dict(self.config, _class_name="UNet2DConditionModel"))

    @classmethod
    def from_pretrained(cls, directory, **_):
        sd, cfg = _load_dir(directory)
        return cls(_config_from_dict(cfg), sd)


class _StandaloneOwner:
    """Engine cache of a net that is used on its own (the stock ControlNet pipeline calls one net alone:
    /root/reference/test_text2image_pretrained_openpose.py:263), outside an EdgeStyleMultiControlNetModel.  The engine
    packs only what that net needs: its own encoder + zero-convs, or -- for a ControlLoRA net -- the UNet's encoder with
    the net's LoRA and zero-convs."""

    def __init__(self, net, dtype=torch.float16, n_text: int = 77):
        self.net, self.dtype, self.n_text = net, dtype, n_text
        self._engines: Dict[tuple, Any] = {}

    def engine(self, rows: int, h: int, w: int):
        from ..engine import DenoiseEngine

        key = (rows, h, w)
        eng = self._engines.get(key)
        if eng is None:
            net = self.net
            if net.uses_lora:
                if net._unet is None:
                    raise RuntimeError("tie_weights(unet) first: the base weights of a ControlLoRA net are the UNet's")
                eng = DenoiseEngine(net.config, net._unet.state_dict(), [net.state_dict()], None, None, rows=rows, h=h, w=w,
                                    dtype=self.dtype, n_text=self.n_text, fuse_lora=True)
            else:
                eng = DenoiseEngine(net.config, None, [], net.state_dict(), None, rows=rows, h=h, w=w, dtype=self.dtype,
                                    n_text=self.n_text)
            self._engines[key] = eng
        return eng

    def embed_engine(self, rows: int, h: int, w: int):
        for (_, hh, ww), eng in self._engines.items():
            if (hh, ww) == (h, w):
                return eng
        return self.engine(rows, h, w)

    def lora_group(self, net) -> Optional[int]:
        return 0 if net.uses_lora else None

    def _single_forward(self, net, sample, timestep, ehs, cond, scale, guess_mode):
        B, _, h, w = sample.shape
        return self.engine(B, h, w).single_controlnet(self.lora_group(net), sample, timestep, ehs, cond, scale, guess_mode)

    def _invalidate_engines(self):
        self._engines.clear()


class CachedControlNetModel:
    """ControlNet whose conditioning embedder is skipped when `controlnet_cond` is already latent-sized
    (controllora.py:199-201) -- the only mode the hot path uses (conds are cached by the pipeline)."""

    uses_lora = False

    def __init__(self, config: C.UNetConfig, state_dict: Mapping[str, torch.Tensor]):
        self.config = C.UNetConfig.from_any(config)
        spec = dict(C.encoder_spec(self.config))
        spec.update(C.controlnet_extra_spec(self.config, with_embedder=False))
        _check_spec(state_dict, spec, type(self).__name__, allow_extra=("controlnet_cond_embedding.",))
        self._sd = OrderedDict(state_dict)
        self._owner = None  # set by EdgeStyleMultiControlNetModel
        self._slots: List[int] = []
        self.controlnet_conditioning_channel_order = "rgb"

    def state_dict(self):
        return self._sd

    def _weights_changed(self):
        """The owning multi-ControlNet caches engines that hold packed device copies of these weights (and captured
        CUDA graphs): drop them, so the next forward runs the new weights -- in the reference the modules are live."""
        if self._owner is not None:
            self._owner._invalidate_engines()

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
        for k, v in state_dict.items():
            if k in self._sd:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise ValueError(f"{k}: shape {tuple(v.shape)} != {tuple(self._sd[k].shape)}")
                self._sd[k] = v
            elif strict:
                raise KeyError(k)
        self._weights_changed()

    # -- reference surface -------------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale: float = 1.0,
                class_labels=None, timestep_cond=None, attention_mask=None, added_cond_kwargs=None,
                cross_attention_kwargs=None, guess_mode: bool = False, return_dict: bool = True
                ) -> Union[ControlNetOutput, Tuple]:
        if self.controlnet_conditioning_channel_order not in ("rgb", "bgr"):  # controllora.py:115-126
            raise ValueError(f"unknown `controlnet_conditioning_channel_order`: {self.controlnet_conditioning_channel_order}")
        for name, val in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                          ("attention_mask", attention_mask), ("added_cond_kwargs", added_cond_kwargs)):
            if val is not None:
                raise NotImplementedError(f"{name} is not used by SD1.5 and not implemented")
        if cross_attention_kwargs and cross_attention_kwargs.get("scale", 1.0) != 1.0:
            raise NotImplementedError('cross_attention_kwargs["scale"] != 1')
        if self._owner is None:  # used on its own: a private engine holding just this net's weights
            self._owner = _StandaloneOwner(self)
        if tuple(controlnet_cond.shape[2:]) != tuple(sample.shape[2:]):  # raw image: run the embedder (:199-201)
            if self.controlnet_conditioning_channel_order == "bgr":
                controlnet_cond = torch.flip(controlnet_cond, dims=[1])
            controlnet_cond = self.preprocess_image(controlnet_cond)
        down, mid = self._owner._single_forward(self, sample, timestep, encoder_hidden_states, controlnet_cond,
                                                float(conditioning_scale), guess_mode)
        if not return_dict:
            return down, mid
        return ControlNetOutput(tuple(down), mid)

    __call__ = forward

    def preprocess_image(self, image, repeats: int = 1, noise=None, generator=None):
        """controllora.py:289-290: the conditioning embedder on a raw control image [n, 3, H, W] -> [n, 320, H/8, W/8].

        Plain ControlNets (openpose): ControlNetConditioningEmbedding, 8 convolutions on the GEMM kernel.
        ControlLoRA nets: VAEControlNetConditioningEmbedding (:38-42) = `conv_vae_out(vae.encode(image).latent_dist
        .sample() * scaling_factor)` on `edgestyle_b200.vae.AutoencoderKL` (set with `set_autoencoder`, :634).
        `repeats` > 1 draws that many independent samples per image from ONE encoder pass (rows ordered like
        `torch.cat([image] * repeats)`): the reference duplicates the image for CFG before embedding it
        (edgestyle_pipeline.py:657-662), so its two CFG rows carry independently sampled embeddings.
        `noise` ([repeats * n, 4, H/8, W/8]) replaces the RNG draw (parity tests); the openpose path ignores both."""
        if self._owner is None:
            self._owner = _StandaloneOwner(self)
        n, _, H, W = image.shape
        if not self.uses_lora:
            emb = self._owner.embed_engine(n, H // 8, W // 8).embed_openpose(image)
            return torch.cat([emb] * repeats) if repeats > 1 else emb
        vae = getattr(self, "autoencoder", None)
        if vae is None:
            raise RuntimeError("ControlLoRAModel.preprocess_image needs the VAE: call set_autoencoder(vae) first "
                               "(controllora.py:634; app.py passes vae= to from_pretrained)")
        dist = vae.encode(image).latent_dist
        if repeats > 1:
            dist = dist.repeat(repeats)
        z = dist.sample(generator=generator, noise=noise, scale=vae.config.scaling_factor)  # :39-40
        # :41 -- conv_vae_out IS this net's conv_in (with its conv LoRA, if any)
        return self._owner.embed_engine(z.shape[0], H // 8, W // 8).embed_vae_latent(z, self._owner.lora_group(self))

    # -- checkpoint format (diffusers layout: <dir>/config.json + <dir>/diffusion_pytorch_model.safetensors) ------
    def _extra_config(self) -> dict:
        return {"_class_name": "ControlNetModel",
                "controlnet_conditioning_channel_order": self.controlnet_conditioning_channel_order}

    def save_pretrained(self, directory, **_):
        _save_dir(directory, self.state_dict(), _config_dict(self.config, **self._extra_config()))

    @classmethod
    def from_pretrained(cls, directory, **_):
        sd, cfg = _load_dir(directory)
        net = cls(_config_from_dict(cfg), sd)
        net.controlnet_conditioning_channel_order = cfg.get("controlnet_conditioning_channel_order", "rgb")
        return net


class ControlLoRAModel(CachedControlNetModel):
    """ControlNet whose conv_in / time_embedding / down_blocks / mid_block ARE the UNet's parameters
    (tie_weights) plus a rank-r LoRA on every Linear under `_skip_layers` (controllora.py:443-450,529-593).

    `state_dict()` holds only LoRA tensors and the non-tied tensors (zero-convs), exactly like :600-606."""

    _skip_layers = _SKIP_LAYERS
    uses_lora = True

    def __init__(self, config: C.UNetConfig, state_dict: Mapping[str, torch.Tensor], lora_linear_rank: int = 4,
                 lora_conv2d_rank: int = 0, unet: Optional[UNet2DConditionModel] = None):
        self.config = C.UNetConfig.from_any(config)
        self.lora_linear_rank, self.lora_conv2d_rank = lora_linear_rank, lora_conv2d_rank
        # lora_conv2d_rank > 0: a LoRAConv2dLayer on every convolution too -- of rank lora_linear_rank (controllora.py:569)
        spec = dict(C.lora_spec(self.config, lora_linear_rank, lora_conv2d_rank))
        spec.update(C.controlnet_extra_spec(self.config, with_embedder=False))
        filtered = OrderedDict((k, v) for k, v in state_dict.items()
                               if k.split(".")[0] not in self._skip_layers or ".lora_layer." in k)
        _check_spec(filtered, spec, "ControlLoRAModel", allow_extra=("controlnet_cond_embedding.",))
        self._sd = filtered
        self._unet = unet
        self._owner = None
        self._slots = []
        self.controlnet_conditioning_channel_order = "rgb"
        self.autoencoder = None

    @property
    def uses_vae(self) -> bool:
        return self.autoencoder is not None

    @classmethod
    def from_unet(cls, unet: UNet2DConditionModel, conditioning_channels: int = 3,
                  controlnet_conditioning_channel_order: str = "rgb",
                  conditioning_embedding_out_channels=(16, 32, 96, 256), lora_linear_rank: int = 4,
                  lora_conv2d_rank: int = 0, autoencoder=None, generator: Optional[torch.Generator] = None):
        """Fresh LoRA (down ~ N(0, 1/r^2), up = 0) and zero zero-convs, tied to `unet` (controllora.py:644-725)."""
        cfg = unet.config
        if conditioning_channels != cfg.conditioning_channels or \
                tuple(conditioning_embedding_out_channels) != tuple(cfg.conditioning_embedding_out_channels):
            raise NotImplementedError("a ControlLoRA net embeds its control image with the VAE (controllora.py:596-598): "
                                      "non-default conditioning_channels / conditioning_embedding_out_channels are unused")
        sd = OrderedDict()
        for k, shape in C.lora_spec(cfg, lora_linear_rank, lora_conv2d_rank).items():
            if k.endswith("down.weight"):
                sd[k] = torch.randn(shape, generator=generator) / lora_linear_rank
            else:
                sd[k] = torch.zeros(shape)
        for k, shape in C.controlnet_extra_spec(cfg, with_embedder=False).items():
            sd[k] = torch.zeros(shape)
        net = cls(cfg, sd, lora_linear_rank, lora_conv2d_rank, unet)
        net.controlnet_conditioning_channel_order = controlnet_conditioning_channel_order
        if autoencoder is not None:  # controllora.py:718-722
            net.set_autoencoder(autoencoder)
        return net

    def tie_weights(self, unet: UNet2DConditionModel):
        changed = unet is not self._unet
        self._unet = unet
        if changed:
            self._weights_changed()

    def set_autoencoder(self, autoencoder):
        self.autoencoder = autoencoder

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
        for k, v in state_dict.items():
            if k in self._sd:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise ValueError(f"{k}: shape {tuple(v.shape)} != {tuple(self._sd[k].shape)}")
                self._sd[k] = v
            elif strict and (k.split(".")[0] not in self._skip_layers) and not k.startswith("controlnet_cond_embedding."):
                raise KeyError(k)
        self._weights_changed()

    def _extra_config(self) -> dict:
        return {"_class_name": "ControlLoRAModel", "lora_linear_rank": self.lora_linear_rank,
                "lora_conv2d_rank": self.lora_conv2d_rank, "uses_vae": self.uses_vae,
                "controlnet_conditioning_channel_order": self.controlnet_conditioning_channel_order}

    @classmethod
    def from_pretrained(cls, directory, unet: Optional[UNet2DConditionModel] = None, **_):
        """Loads only LoRA + non-tied tensors (what :600-606 saves); call tie_weights(unet) afterwards (app.py:95-97)."""
        sd, cfg = _load_dir(directory)
        net = cls(_config_from_dict(cfg), sd, lora_linear_rank=cfg.get("lora_linear_rank", 4),
                  lora_conv2d_rank=cfg.get("lora_conv2d_rank", 0), unet=unet)
        net.controlnet_conditioning_channel_order = cfg.get("controlnet_conditioning_channel_order", "rgb")
        return net

    def fused_state_dict(self, lora_scale: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
        """Full diffusers-layout ControlNet state dict of this net with the LoRA folded in: every tied base tensor of the
        UNet encoder, `W + lora_scale * up @ down` for each LoRA'd Linear (diffusers `LoRACompatibleLinear._fuse_lora`,
        reached from controllora.py:728-737), plus the net's own tensors (zero-convs).  fp32 host math; the UNet is not
        modified."""
        if self._unet is None:
            raise RuntimeError("tie_weights(unet) first: the base weights of a ControlLoRA net are the UNet's (controllora.py:623-632)")
        usd = self._unet.state_dict()
        out = OrderedDict()
        for k in C.encoder_spec(self.config):
            w = usd[k]
            if k.endswith(".weight"):
                base = k[:-len(".weight")]
                dn, up = self._sd.get(f"{base}.lora_layer.down.weight"), self._sd.get(f"{base}.lora_layer.up.weight")
                if dn is not None and up is not None:  # Linear, or LoRACompatibleConv._fuse_lora (flatten(1) on both)
                    upd = (up.float().flatten(1) @ dn.float().flatten(1)).reshape(w.shape)
                    w = (w.float() + lora_scale * upd).to(w.dtype)
            out[k] = w
        for k, v in self._sd.items():  # conv_vae_out is an alias of the (tied) conv_in module (controllora.py:36)
            if ".lora_layer." not in k and k not in out and not k.startswith("controlnet_cond_embedding.conv_vae_out."):
                out[k] = v
        return out

    def fuse(self) -> "FusedControlLoRAModel":
        """controllora.py:739-777: a plain ControlNet holding W + up @ down (e.g. to export one self-contained
        checkpoint).  Unlike the reference -- whose `fuse_lora` writes through the tied Parameters and so also rewrites
        the UNet (and stacks both nets' updates when two ControlLoRAs share it) -- the UNet is left untouched."""
        net = FusedControlLoRAModel(self.config, self.fused_state_dict(1.0))
        net.controlnet_conditioning_channel_order = self.controlnet_conditioning_channel_order
        if self.autoencoder is not None:
            net.autoencoder = self.autoencoder
        return net

    def fuse_lora(self, lora_scale: float = 1.0, safe_fusing: bool = False):
        raise NotImplementedError(
            "in-place fusing would write through the tied UNet weights (the reference's fuse_lora does, corrupting the "
            "UNet); use fuse() / fused_state_dict() for a fused copy, or EdgeStyleMultiControlNetModel.fuse() to run "
            "the engine on fused weight copies (its default)")


class FusedControlLoRAModel(CachedControlNetModel):
    """controllora.py:292-375: a plain ControlNet whose weights already contain the LoRA update (`ControlLoRAModel
    .fuse()`).  A weight container in diffusers ControlNet layout (save_pretrained / from_pretrained); inside an
    EdgeStyleMultiControlNetModel the fused path is selected with `multi.fuse()` instead, which keeps the tied base
    pass and the cached conditioning embeddings."""

    uses_lora = False

    def preprocess_image(self, image, **_):
        raise NotImplementedError("FusedControlLoRAModel is a weight container here: embed control images with the "
                                  "ControlLoRAModel it came from (its embedder is the VAE, not the openpose convs)")

    def _extra_config(self) -> dict:
        return {"_class_name": "FusedControlLoRAModel", "uses_vae": True}
