"""AutoencoderKL (SD1.5 VAE) on the sm_100a kernels -- the per-call stages either side of the denoise loop
(SURVEY.md 8(f) row N2).

Reference call sites: `VAEControlNetConditioningEmbedding.forward`, /root/reference/model/controllora.py:38-42
(`autoencoder.encode(img).latent_dist.sample() * scaling_factor`, cached once per call by
/root/reference/model/edgestyle_pipeline.py:660-662) and the final `vae.decode(latents / scaling_factor)`,
edgestyle_pipeline.py:552-557.  The class keeps the diffusers surface those lines use (`encode(x).latent_dist.sample()
/ .mode()`, `decode(z).sample`, `config.scaling_factor`) and takes a diffusers-layout `vae/` state dict.

encode / decode / sample run channels-last on the library's kernels, with no torch arithmetic (the `logvar` / `std`
accessors of the distribution, which the reference path never reads, are plain torch views of the moments):
  * 3x3 convolutions: es_gemm implicit GEMM (tcgen05); the stride-2 `Downsample2D(padding=0)` convs pad (0, 1, 0, 1):
    es_im2col3x3_pad + flat es_gemm; `Upsample2D`: es_upsample2x + es_gemm.
  * GroupNorm(32, eps 1e-6) + SiLU: the statistics come from the epilogue of the GEMM that produced the tensor
    (EsGemm.gn_ws), so only es_groupnorm_apply runs; the residual add rides in the conv2 epilogue and a resnet's 1x1
    shortcut is a second K source of its conv2 (EsGemm.a2 / b2).
  * mid-block attention: ONE head of C = 512 channels over h*w tokens -- wider than es_attention's TMEM budget (192),
    so S = Q K^T (es_gemm, fp32 out, scale as `alpha`), P = es_softmax_rows(S), O = P V (es_gemm against V^T, which
    a GEMM with swapped operands produces directly: V^T = W_v X^T).  The V bias is folded into the output bias
    (rows of P sum to 1: P (V + 1 b_v^T) W_o^T + b_o = P V W_o^T + (W_o b_v + b_o)).
  * encoder tail: `quant_conv` (1x1) is folded into `conv_out` on the host (exact: a 1x1 after a conv is a conv).
  * DiagonalGaussianDistribution.sample: es_gaussian_sample.
There is no CPU path: construction without CUDA raises.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass, fields
from typing import Dict, List, Mapping, Optional, Tuple

import torch

from . import ops
from .ext import EdgeStyleNativeError


@dataclass
class VaeConfig:
    """SD1.5 `vae/config.json` subset."""

    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    norm_eps: float = 1e-6

    @classmethod
    def from_any(cls, cfg) -> "VaeConfig":
        if isinstance(cfg, cls):
            return cfg
        if cfg is None:
            return cls()
        get = cfg.get if isinstance(cfg, Mapping) else (lambda k, d=None: getattr(cfg, k, d))
        kw = {}
        for f in fields(cls):
            v = get(f.name, None)
            if v is not None:
                kw[f.name] = tuple(v) if f.name == "block_out_channels" else v
        return cls(**kw)


def vae_spec(cfg: VaeConfig) -> "OrderedDict[str, tuple]":
    """diffusers AutoencoderKL state-dict names and shapes (checked against the oracle in tests/test_host_cpu.py)."""
    ch = list(cfg.block_out_channels)
    sp: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(name, cout, cin, k):
        sp[name + ".weight"] = (cout, cin, k, k)
        sp[name + ".bias"] = (cout,)

    def norm(name, c):
        sp[name + ".weight"] = (c,)
        sp[name + ".bias"] = (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cout, cin, 1)

    def mid(name, c):
        a = name + ".attentions.0"
        norm(a + ".group_norm", c)
        for p in ("to_q", "to_k", "to_v", "to_out.0"):
            sp[f"{a}.{p}.weight"] = (c, c)
            sp[f"{a}.{p}.bias"] = (c,)
        resnet(name + ".resnets.0", c, c)
        resnet(name + ".resnets.1", c, c)

    conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i < len(ch) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        cin = cout
    mid("encoder.mid_block", ch[-1])
    norm("encoder.conv_norm_out", ch[-1])
    conv("encoder.conv_out", 2 * cfg.latent_channels, ch[-1], 3)
    rev = ch[::-1]
    conv("decoder.conv_in", rev[0], cfg.latent_channels, 3)
    mid("decoder.mid_block", rev[0])
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i < len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        cin = cout
    norm("decoder.conv_norm_out", ch[0])
    conv("decoder.conv_out", cfg.out_channels, ch[0], 3)
    conv("quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    return sp


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


class DiagonalGaussianDistribution:
    """Device-side mirror of diffusers' DiagonalGaussianDistribution over moments kept channels-last
    ([n*h*w, ld] fp32, columns = mean | logvar)."""

    def __init__(self, moments: torch.Tensor, n: int, latent_channels: int, h: int, w: int):
        self._m, self._n, self._L, self._h, self._w = moments, n, latent_channels, h, w

    def _draw(self, noise: Optional[torch.Tensor], scale: float) -> torch.Tensor:
        out = torch.empty(self._n, self._L, self._h, self._w, device=self._m.device, dtype=torch.float32)
        return ops.gaussian_sample(self._m, noise, out, scale)

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None,
               scale: float = 1.0) -> torch.Tensor:
        """mean + std * N(0, 1) (`noise` may be handed in, NCHW fp32, for reproducible parity tests); `scale`
        multiplies the result inside the same kernel (controllora.py:40)."""
        if noise is None:  # diffusers randn_tensor: drawn on the distribution's device from `generator` / global RNG
            noise = torch.randn(self._n, self._L, self._h, self._w, generator=generator, device=self._m.device,
                                dtype=torch.float32)
        else:
            noise = noise.to(device=self._m.device, dtype=torch.float32).contiguous()
        return self._draw(noise, scale)

    def mode(self, scale: float = 1.0) -> torch.Tensor:
        return self._draw(None, scale)

    def repeat(self, repeats: int) -> "DiagonalGaussianDistribution":
        """The distribution of `torch.cat([x] * repeats)` without re-running the encoder (a row copy of the moments)."""
        return DiagonalGaussianDistribution(self._m.repeat(repeats, 1), self._n * repeats, self._L, self._h, self._w)

    @property
    def mean(self) -> torch.Tensor:
        return self.mode()

    @property
    def logvar(self) -> torch.Tensor:
        L = self._L
        lv = self._m[:, L:2 * L].reshape(self._n, self._h, self._w, L).permute(0, 3, 1, 2)
        return lv.clamp(-30.0, 20.0).contiguous()

    @property
    def std(self) -> torch.Tensor:
        return torch.exp(0.5 * self.logvar)


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


@dataclass
class DecoderOutput:
    sample: torch.Tensor


@dataclass
class _Conv3:
    w: torch.Tensor      # [cout_rows, 9 * cin_pad] (tap-major: column = tap * cin_pad + channel)
    b: torch.Tensor      # fp32 [cout_rows]
    cin_pad: int
    cout: int


@dataclass
class _Res:
    n1: Tuple[torch.Tensor, torch.Tensor]
    c1: _Conv3
    n2: Tuple[torch.Tensor, torch.Tensor]
    c2: _Conv3
    sc: Optional[torch.Tensor]  # 1x1 shortcut weight [cout, cin]: second K source of conv2 (its bias rides in c2.b)
    cin: int
    cout: int


@dataclass
class _Attn:
    gn: Tuple[torch.Tensor, torch.Tensor]
    wq: torch.Tensor
    bq: torch.Tensor
    wk: torch.Tensor
    bk: torch.Tensor
    wv: torch.Tensor
    wo: torch.Tensor
    bo: torch.Tensor  # b_o + W_o b_v


class AutoencoderKL:
    """Drop-in for the diffusers AutoencoderKL calls of the reference (see module docstring)."""

    def __init__(self, config=None, state_dict: Optional[Mapping[str, torch.Tensor]] = None, dtype=torch.float16,
                 device="cuda"):
        self.config = VaeConfig.from_any(config)
        cfg = self.config
        if state_dict is None:
            raise ValueError("state_dict (diffusers `vae/` layout) is required")
        spec = vae_spec(cfg)
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"AutoencoderKL: missing keys {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k, shp in spec.items():
            got = tuple(state_dict[k].shape)
            if got != tuple(shp) and got != tuple(shp) + (1, 1):  # old checkpoints store the attention Linears as 1x1 convs
                raise ValueError(f"AutoencoderKL: {k} has shape {got}, expected {tuple(shp)}")
        for c in cfg.block_out_channels:
            if c % 8 or c % cfg.norm_num_groups:
                raise ValueError(f"channel count {c} unsupported (needs %8 and %norm_num_groups)")
        if 2 * cfg.latent_channels > 16 or cfg.out_channels > 16:
            raise ValueError("latent / output channel count unsupported")
        if not torch.cuda.is_available():
            raise EdgeStyleNativeError("edgestyle_b200.vae.AutoencoderKL needs a CUDA device (there is no CPU path)")
        self.dev = torch.device(device)
        self.dtype = dtype
        self._sd = OrderedDict((k, state_dict[k]) for k in spec)
        self._bufs: Dict[str, torch.Tensor] = {}
        self._stats_of: Dict[tuple, torch.Tensor] = {}  # tensor -> statistics slot filled by its producing GEMM
        self._ws_next = 0
        self.fuse_gn_stats = os.environ.get("ES_VAE_FUSE_GN", "1") != "0"
        ops.set_gemm_workspace(256 << 20, self.dev)
        self._pack()

    def state_dict(self):
        return self._sd

    # ------------------------------------------------------------------------------------ packing
    def _mat(self, t):
        return t.to(device=self.dev, dtype=self.dtype).contiguous()

    def _f32(self, t):
        return t.to(device=self.dev, dtype=torch.float32).contiguous()

    def _w(self, name):
        return self._sd[name].detach().float().cpu()

    def _conv3(self, name, w=None, b=None, pad_rows: int = 0) -> _Conv3:
        w = self._w(name + ".weight") if w is None else w
        b = self._w(name + ".bias") if b is None else b
        cout, cin = w.shape[:2]
        cin_pad = _pad8(cin)
        rows = max(cout, pad_rows)
        wp = torch.zeros(rows, 3, 3, cin_pad)
        wp[:cout, ..., :cin] = w.permute(0, 2, 3, 1)
        bp = torch.zeros(rows)
        bp[:cout] = b
        return _Conv3(self._mat(wp.reshape(rows, 9 * cin_pad)), self._f32(bp), cin_pad, cout)

    def _norm(self, name):
        return self._f32(self._w(name + ".weight")), self._f32(self._w(name + ".bias"))

    def _res(self, name, cin, cout) -> _Res:
        sc, b2 = None, self._w(name + ".conv2.bias")
        if cin != cout:  # conv_shortcut(x) + conv2(h): one accumulator, two K sources (EsGemm.a2 / b2)
            sc = self._mat(self._w(name + ".conv_shortcut.weight").reshape(cout, cin))
            b2 = b2 + self._w(name + ".conv_shortcut.bias")
        return _Res(self._norm(name + ".norm1"), self._conv3(name + ".conv1"), self._norm(name + ".norm2"),
                    self._conv3(name + ".conv2", b=b2), sc, cin, cout)

    def _attn(self, name, c) -> _Attn:
        lin = lambda p: self._w(f"{name}.{p}.weight").reshape(c, c)
        wo, bv = lin("to_out.0"), self._w(f"{name}.to_v.bias")
        return _Attn(self._norm(name + ".group_norm"), self._mat(lin("to_q")), self._f32(self._w(f"{name}.to_q.bias")),
                     self._mat(lin("to_k")), self._f32(self._w(f"{name}.to_k.bias")), self._mat(lin("to_v")),
                     self._mat(wo), self._f32(self._w(f"{name}.to_out.0.bias") + wo @ bv))

    def _pack(self):
        cfg = self.config
        ch = list(cfg.block_out_channels)
        L = cfg.latent_channels
        # ---- encoder
        self.e_conv_in = self._conv3("encoder.conv_in")
        self.e_down: List[Tuple[List[_Res], Optional[Tuple[torch.Tensor, torch.Tensor]]]] = []
        cin = ch[0]
        for i, cout in enumerate(ch):
            res = [self._res(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
                   for j in range(cfg.layers_per_block)]
            ds = None
            if i < len(ch) - 1:
                k = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                ds = (self._mat(self._w(k + ".weight").permute(0, 2, 3, 1).reshape(cout, 9 * cout)),
                      self._f32(self._w(k + ".bias")))
            self.e_down.append((res, ds))
            cin = cout
        c = ch[-1]
        self.e_mid = (self._res("encoder.mid_block.resnets.0", c, c), self._attn("encoder.mid_block.attentions.0", c),
                      self._res("encoder.mid_block.resnets.1", c, c))
        self.e_norm_out = self._norm("encoder.conv_norm_out")
        # quant_conv (1x1, 2L -> 2L) folded into conv_out: W' = Wq Wco, b' = Wq bco + bq
        wq = self._w("quant_conv.weight").reshape(2 * L, 2 * L)
        wco, bco = self._w("encoder.conv_out.weight"), self._w("encoder.conv_out.bias")
        self.e_conv_out = self._conv3(None, torch.einsum("om,mikl->oikl", wq, wco),
                                      wq @ bco + self._w("quant_conv.bias"), pad_rows=16)
        # ---- decoder
        wpq = torch.zeros(8, 64)
        wpq[:L, :L] = self._w("post_quant_conv.weight").reshape(L, L)
        bpq = torch.zeros(8)
        bpq[:L] = self._w("post_quant_conv.bias")
        self.d_post_quant = (self._mat(wpq), self._f32(bpq))
        rev = ch[::-1]
        self.d_conv_in = self._conv3("decoder.conv_in")
        c = rev[0]
        self.d_mid = (self._res("decoder.mid_block.resnets.0", c, c), self._attn("decoder.mid_block.attentions.0", c),
                      self._res("decoder.mid_block.resnets.1", c, c))
        self.d_up: List[Tuple[List[_Res], Optional[_Conv3]]] = []
        cin = rev[0]
        for i, cout in enumerate(rev):
            res = [self._res(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
                   for j in range(cfg.layers_per_block + 1)]
            us = self._conv3(f"decoder.up_blocks.{i}.upsamplers.0.conv") if i < len(ch) - 1 else None
            self.d_up.append((res, us))
            cin = cout
        self.d_norm_out = self._norm("decoder.conv_norm_out")
        self.d_conv_out = self._conv3("decoder.conv_out", pad_rows=16)

    # ------------------------------------------------------------------------------------ building blocks
    def buf(self, name: str, rows: int, cols: int, dtype=None) -> torch.Tensor:
        dtype = dtype or self.dtype
        key = f"{name}:{rows}x{cols}:{dtype}"
        t = self._bufs.get(key)
        if t is None:
            t = torch.zeros(rows, cols, device=self.dev, dtype=dtype)
            self._bufs[key] = t
        return t

    def release_buffers(self):
        """Drop the activation scratch (a 512x512 image keeps ~0.5 GB of it alive between calls)."""
        self._bufs.clear()
        self._stats_of.clear()

    # Scratch is keyed by role and shape, not by layer: a GroupNorm output is consumed by the very next launch, conv1's
    # output by the next GroupNorm, and the resnet outputs ping-pong between two slots (the input of a resnet is the
    # residual of its conv2, so it must outlive it).
    def _stats_kw(self, out, n, hw, flat: bool) -> dict:
        """GroupNorm statistics of `out` accumulated by the GEMM that produces it (EsGemm.gn_ws): the GroupNorm that
        consumes `out` then only applies.  Needs whole 32-row groups per image and a dense 16-bit output."""
        G = self.config.norm_num_groups
        # below 8 channels per group the epilogue falls back to per-chunk warp reductions + global atomics on a handful
        # of (image, group) addresses: measured 10.2 vs 6.2 ms for the 512x512 encoder (profiles/README.md), so those
        # layers (the 128-channel level) keep the separate statistics pass
        if (not self.fuse_gn_stats or hw % 32 or out.shape[1] % G or out.shape[1] // G < 8
                or out.stride(0) != out.shape[1]):
            return {}
        self._ws_next = (self._ws_next + 1) % 3
        ws = self.buf(f"gn.ws{self._ws_next}", n, 2 * G, torch.float32)
        ws.zero_()
        self._stats_of[(out.data_ptr(), out.shape[0], out.shape[1])] = ws
        kw = dict(gn_ws=ws, gn_groups=G)
        if flat:
            kw["rows_per_img"] = hw
        return kw

    def _gn(self, x, gb, n, hw, silu):
        out = self.buf("gn", x.shape[0], x.shape[1])
        G, eps = self.config.norm_num_groups, self.config.norm_eps
        ws = self._stats_of.pop((x.data_ptr(), x.shape[0], x.shape[1]), None)
        if ws is not None:
            return ops.groupnorm(x, out, gb[0], gb[1], ws, n, hw, G, eps, silu, stats_ready=True)
        return ops.groupnorm(x, out, gb[0], gb[1], self.buf("gn.ws", n, 2 * G, torch.float32), n, hw, G, eps, silu)

    def _conv(self, tag, x, cv: _Conv3, n, H, W, residual=None, a2=None, b2=None, stats: bool = True):
        out = self.buf(f"conv.{tag}", n * H * W, cv.cout)
        kw = self._stats_kw(out, n, H * W, False) if stats else {}
        return ops.gemm(x, cv.w, cv.cout, out=out, taps=9, whn=(W, H, n), bias=cv.b, c1=cv.cin_pad, residual=residual,
                        a2=a2, b2=b2, **kw)

    def _resnet(self, slot: int, x, r: _Res, n, H, W):
        h = self._gn(x, r.n1, n, H * W, True)
        h = self._conv("r1", h, r.c1, n, H, W)
        h = self._gn(h, r.n2, n, H * W, True)
        if r.sc is not None:
            return self._conv(f"r2.{slot % 2}", h, r.c2, n, H, W, a2=x, b2=r.sc)
        return self._conv(f"r2.{slot % 2}", h, r.c2, n, H, W, residual=x)

    def _attention(self, tag, x, a: _Attn, n, hw):
        C = x.shape[1]
        hn = self._gn(x, a.gn, n, hw, False)
        q = ops.gemm(hn, a.wq, C, out=self.buf(f"{tag}.q", n * hw, C), bias=a.bq)
        k = ops.gemm(hn, a.wk, C, out=self.buf(f"{tag}.k", n * hw, C), bias=a.bk)
        o = self.buf(f"{tag}.o", n * hw, C)
        s = self.buf(f"{tag}.s", hw, hw, torch.float32)
        p = self.buf(f"{tag}.p", hw, hw)
        vt = self.buf(f"{tag}.vt", C, hw)
        for i in range(n):
            rows = slice(i * hw, (i + 1) * hw)
            ops.gemm(q[rows], k[rows], hw, out=s, alpha=C ** -0.5)     # S = Q K^T / sqrt(C)
            ops.softmax_rows(s, p)
            ops.gemm(a.wv, hn[rows], hw, out=vt)                        # V^T = W_v X^T (bias folded into bo)
            ops.gemm(p, vt, C, out=o[rows])                             # O = P V
        out = self.buf(f"{tag}.out", n * hw, C)
        return ops.gemm(o, a.wo, C, out=out, bias=a.bo, residual=x, **self._stats_kw(out, n, hw, True))

    def _mid(self, x, mid, n, H, W):
        x = self._resnet(0, x, mid[0], n, H, W)
        x = self._attention("att", x, mid[1], n, H * W)   # own output buffer: resnet slot 0 is its residual
        return self._resnet(1, x, mid[2], n, H, W)

    # ------------------------------------------------------------------------------------ public
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x: [n, 3, H, W] fp32 NCHW in [-1, 1] -> latent_dist over [n, L, H/8, W/8]."""
        cfg = self.config
        n, cin, H, W = x.shape
        down = 2 ** (len(cfg.block_out_channels) - 1)
        if cin != cfg.in_channels or H % down or W % down:
            raise ValueError(f"encode: expected [n, {cfg.in_channels}, H, W] with H, W multiples of {down}")
        t = f"e{n}x{H}x{W}"
        self._stats_of.clear()
        a = self.buf(f"{t}.in", n * H * W, self.e_conv_in.cin_pad)
        ops.nchw_to_nhwc(x.to(device=self.dev, dtype=torch.float32).contiguous(), a)
        a = self._conv(f"{t}.in", a, self.e_conv_in, n, H, W)
        for i, (res, ds) in enumerate(self.e_down):
            for j, r in enumerate(res):
                a = self._resnet(j, a, r, n, H, W)
            if ds is not None:  # Downsample2D(padding=0): zero pad right/bottom, 3x3 stride 2
                c = a.shape[1]
                Ho, Wo = H // 2, W // 2
                col = self.buf(f"{t}.col{i}", n * Ho * Wo, 9 * c)
                ops.im2col3x3_pad(a, col, n, H, W, c, 2, 0, 1)
                dso = self.buf(f"{t}.ds{i}", n * Ho * Wo, c)
                a = ops.gemm(col, ds[0], c, out=dso, bias=ds[1], **self._stats_kw(dso, n, Ho * Wo, True))
                H, W = Ho, Wo
        a = self._mid(a, self.e_mid, n, H, W)
        a = self._gn(a, self.e_norm_out, n, H * W, True)
        mom = self.buf(f"{t}.moments", n * H * W, 16, torch.float32)
        ops.gemm(a, self.e_conv_out.w, 2 * cfg.latent_channels, out=mom, taps=9, whn=(W, H, n),
                 bias=self.e_conv_out.b, c1=self.e_conv_out.cin_pad, block_n=32)
        dist = DiagonalGaussianDistribution(mom.clone(), n, cfg.latent_channels, H, W)
        return AutoencoderKLOutput(dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """z: [n, L, h, w] fp32 NCHW (already divided by scaling_factor) -> image [n, 3, 8h, 8w] fp32 NCHW."""
        cfg = self.config
        n, L, H, W = z.shape
        if L != cfg.latent_channels:
            raise ValueError(f"decode: expected {cfg.latent_channels} latent channels, got {L}")
        t = f"d{n}x{H}x{W}"
        self._stats_of.clear()
        zin = self.buf(f"{t}.z", n * H * W, 64)
        ops.nchw_to_nhwc(z.to(device=self.dev, dtype=torch.float32).contiguous(), zin)
        a = ops.gemm(zin, self.d_post_quant[0], 8, out=self.buf(f"{t}.pq", n * H * W, 8), bias=self.d_post_quant[1])
        a = self._conv(f"{t}.in", a, self.d_conv_in, n, H, W)
        a = self._mid(a, self.d_mid, n, H, W)
        for i, (res, us) in enumerate(self.d_up):
            for j, r in enumerate(res):
                a = self._resnet(j, a, r, n, H, W)
            if us is not None:  # Upsample2D: nearest x2, 3x3 conv
                up = self.buf(f"{t}.up{i}", n * 4 * H * W, a.shape[1])
                ops.upsample2x(a, up, n, H, W)
                H, W = 2 * H, 2 * W
                a = self._conv(f"{t}.us{i}", up, us, n, H, W)
        a = self._gn(a, self.d_norm_out, n, H * W, True)
        o = self.buf(f"{t}.img", n * H * W, 16, torch.float32)
        ops.gemm(a, self.d_conv_out.w, cfg.out_channels, out=o, taps=9, whn=(W, H, n), bias=self.d_conv_out.b,
                 c1=self.d_conv_out.cin_pad, block_n=32)
        img = torch.empty(n, cfg.out_channels, H, W, device=self.dev, dtype=torch.float32)
        # [n*H*W, 16] fp32 (first out_channels valid) -> NCHW: a strided copy, no arithmetic
        img.copy_(o.view(n, H * W, 16)[:, :, :cfg.out_channels].permute(0, 2, 1).reshape(img.shape))
        return DecoderOutput(img) if return_dict else (img,)

    # ------------------------------------------------------------------------------------ checkpoint format
    def save_pretrained(self, directory, **_):
        from .model.controllora import _save_dir

        cfg = {f.name: (list(getattr(self.config, f.name)) if f.name == "block_out_channels"
                        else getattr(self.config, f.name)) for f in fields(VaeConfig)}
        cfg["_class_name"] = "AutoencoderKL"
        _save_dir(directory, self.state_dict(), cfg)

    @classmethod
    def from_pretrained(cls, directory, torch_dtype=None, device="cuda", **_):
        from .model.controllora import _load_dir

        sd, cfg = _load_dir(directory)
        return cls(cfg, sd, dtype=torch_dtype or torch.float16, device=device)
