"""DenoiseEngine: weight packing + the static launch schedule of one EdgeStyle denoise step on one B200.

What the reference does per step (/root/reference/model/edgestyle_pipeline.py:434-543):
six ControlNet encoders run one after another, their residuals are interleaved and merged, then the UNet
runs.  Here the seven encoder traversals collapse into TWO batched passes because only three distinct
weight sets exist (F6: pattern [lora0, pose, lora1, pose, lora1, pose], ControlLoRA base weights tied
to the UNet):

  base pass : UNet-encoder rows | agnostic rows | clothes(cond 2) rows | clothes(cond 4) rows
              = 4*B images through the UNet's conv/linear weights; the three row segments differ only
              in the rank-r LoRA update of each Linear, executed as extra K-blocks of the same
              tcgen05 accumulator (es_gemm source 2) selected per row segment.
  pose pass : 3*B images through the openpose ControlNet weights.
  decoder   : B images (UNet up blocks); `torch.cat([x, skip])` never materialises separately: producers
              write straight into column slices of the concat buffer, and the EdgeStyle merge kernel adds the
              ControlNet residual to the UNet skip while doing so.

Everything is enqueued on the current CUDA stream through the C-ABI (edgestyle_b200.ops); with
`use_graph=True` one step is captured into a CUDA graph and replayed.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import config as C
from . import ops
from .ext import ACT_GEGLU, ACT_SILU

SUPPORTED_PATTERN = (0, None, 1, None, 1, None)


def pack_merge_block(sd: Dict[str, torch.Tensor], Cc: int, h: int, w: int, dtype, device) -> Dict[str, torch.Tensor]:
    """Repack one ControlNetBlock (/root/reference/model/edgestyle_multicontrolnet.py:23-63) for es_merge_phase.

    The reference's interleaved channel index is c*6 + net; first_conv group g = c*3 + p consumes nets
    (2p, 2p+1) of channel c.  LayerNorm affine [3C, H, W] / [C, H, W] become channels-last [hw, 3, C] / [hw, C].
    """
    hw = h * w
    # (permute + cast on the host, one copy to the device: no packing kernels on the GPU)
    f32 = lambda t: t.detach().to("cpu", torch.float32).contiguous().to(device)
    dd = lambda t: t.detach().to("cpu", dtype).contiguous().to(device)
    return {
        "w1": f32(sd["first_conv.weight"].reshape(Cc, 3, 2).permute(1, 2, 0)),   # [3 pairs][2 nets][C]
        "b1": f32(sd["first_conv.bias"].reshape(Cc, 3).permute(1, 0)),            # [3][C]
        "g1": dd(sd["first_normalization.weight"].reshape(Cc, 3, hw).permute(2, 1, 0)),
        "be1": dd(sd["first_normalization.bias"].reshape(Cc, 3, hw).permute(2, 1, 0)),
        "w2": f32(sd["second_conv.weight"].reshape(Cc, 3).permute(1, 0)),         # [3][C]
        "b2": f32(sd["second_conv.bias"].reshape(Cc)),
        "g2": dd(sd["second_normalization.weight"].reshape(Cc, hw).permute(1, 0)),
        "be2": dd(sd["second_normalization.bias"].reshape(Cc, hw).permute(1, 0)),
        "w3": f32(sd["third_conv.weight"].reshape(Cc)),
        "b3": f32(sd["third_conv.bias"].reshape(Cc)),
    }


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _geglu_block_n(n: int) -> int:
    # 160 first: the CTA-pair kernel (block_n 320) reads the value/gate permutation in sub-tiles of 160 columns, so
    # both kernels can run the same packed weights
    for bn in (160, 256, 128, 64, 32):
        if n % bn == 0:
            return bn
    raise ValueError(f"GEGLU width {n} is not a multiple of 32")


@dataclass
class Lin:
    """A packed Linear / 1x1 conv: w [n, k] (dtype), bias fp32 [n] or None, optional stacked LoRA."""

    w: torch.Tensor
    bias: Optional[torch.Tensor]
    n: int
    down: Optional[torch.Tensor] = None  # [G * rp, k]
    up: Optional[torch.Tensor] = None    # [G * n, rp]
    rp: int = 0
    block_n: int = 0
    groups: int = 1  # > 1: `w` / `bias` hold (1 + G) copies stacked along N: [W; W + up_1 down_1; ...] (LoRA fused)
    colsum: Optional[torch.Tensor] = None  # folded LayerNorm: sum_k (W * gamma)[n, k] per stacked row (fp32)


@dataclass
class Res:
    n1g: torch.Tensor
    n1b: torch.Tensor
    w1: torch.Tensor
    b1: torch.Tensor
    n2g: torch.Tensor
    n2b: torch.Tensor
    w2: torch.Tensor
    b2: torch.Tensor  # conv2 bias (+ shortcut bias when fused)
    wsc: Optional[torch.Tensor]
    cin: int
    cout: int
    temb_off: int = 0
    # conv LoRA (lora_conv2d_rank > 0, controllora.py:561-575).  Fused mode: w1 / w2 / wsc / b1 / b2 hold `groups`
    # copies stacked along N ([W; W + up_1 down_1; ...]) and an image segment selects its copy.  Unfused mode: the
    # rank-r update stays a K-extension: lora[k] = (down [G * rp, taps * cin], up [G * cout, rp], rp) for
    # k in ("conv1", "conv2", "conv_shortcut")
    groups: int = 1
    lora: Optional[Dict[str, Tuple[torch.Tensor, torch.Tensor, int]]] = None
    bsc: Optional[torch.Tensor] = None  # shortcut bias on its own (only used when the shortcut runs as its own GEMM)


@dataclass
class ConvW:
    """A packed stand-alone convolution (conv_in as im2col GEMM, down-sampler): w [groups * cout, K], bias fp32
    [groups * cout]; unfused conv LoRA as (down [G * rp, K], up [G * cout, rp], rp)."""

    w: torch.Tensor
    bias: torch.Tensor
    cout: int
    groups: int = 1
    lora: Optional[Tuple[torch.Tensor, torch.Tensor, int]] = None


@dataclass
class Tfm:
    ng: torch.Tensor
    nb: torch.Tensor
    proj_in: Lin
    ln1: Tuple[torch.Tensor, torch.Tensor]
    qkv: Lin
    o1: Lin
    ln2: Tuple[torch.Tensor, torch.Tensor]
    q2: Lin
    kv2: Lin
    o2: Lin
    ln3: Tuple[torch.Tensor, torch.Tensor]
    ff1: Lin
    ff2: Lin
    proj_out: Lin
    c: int
    uid: str = ""


class _Packer:
    def __init__(self, sd, loras: Sequence[Dict[str, torch.Tensor]], dtype, device, fuse_lora: bool = True,
                 fold_ln: bool = False):
        self.sd, self.loras, self.dtype, self.device = sd, list(loras), dtype, device
        self.fuse_lora = fuse_lora
        self.fold_ln = fold_ln

    def f32(self, k):
        # (cast on the host, then one copy: packing issues no conversion kernels on the device)
        return self.sd[k].detach().to("cpu", torch.float32).contiguous().to(self.device)

    def conv3(self, k):
        w = self.sd[k]  # [cout, cin, 3, 3] -> [cout, 9*cin] with column = tap*cin + c
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(device=self.device, dtype=self.dtype).contiguous()

    def mat(self, t):
        return t.detach().to("cpu", self.dtype).contiguous().to(self.device)

    def _lora_parts(self, name, k_in):
        """[(down [rp,k], up [n,rp])] per LoRA group, rank zero-padded to a multiple of 8; None if absent."""
        parts = []
        for sd in self.loras:
            dk, uk = f"{name}.lora_layer.down.weight", f"{name}.lora_layer.up.weight"
            if dk not in sd:
                return None
            d, u = sd[dk].float(), sd[uk].float()
            d, u = d.reshape(d.shape[0], -1), u.reshape(u.shape[0], -1)  # 1x1 conv LoRA (proj_in / proj_out) -> 2-D
            r = d.shape[0]
            rp = _pad8(r)
            dp = torch.zeros(rp, k_in)
            dp[:r] = d.cpu()
            upad = torch.zeros(u.shape[0], rp)
            upad[:, :r] = u.cpu()
            parts.append((dp, upad))
        return parts

    def lin(self, names: Sequence[str], bias_names: Sequence[Optional[str]] = (), perm: Optional[torch.Tensor] = None,
            block_n: int = 0, ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Lin:
        """Stack one or more Linear / 1x1-conv weights along N (fused QKV / KV); LoRA ups become block-diagonal.

        ln = (gamma, beta) of the LayerNorm that feeds this projection: folded into the weights (W * gamma, bias +
        W beta, column sums kept for the mean term) so the GEMM consumes the un-normalised hidden state."""
        ws = [self.sd[n + ".weight"].reshape(self.sd[n + ".weight"].shape[0], -1).float().cpu() for n in names]
        k_in = ws[0].shape[1]
        w = torch.cat(ws, 0)
        n_tot = w.shape[0]
        bias = None
        if bias_names and any(b is not None for b in bias_names):
            bias = torch.cat([self.sd[b].float().cpu() if b is not None else torch.zeros(wi.shape[0])
                              for b, wi in zip(bias_names, ws)])
        down = up = None
        rp_tot = 0
        if self.loras:
            per_name = [self._lora_parts(n, k_in) for n in names]
            if all(p is not None for p in per_name):
                G = len(self.loras)
                rp = per_name[0][0][0].shape[0]
                rp_tot = rp * len(names)
                downs, ups = [], []
                for g in range(G):
                    downs.append(torch.cat([per_name[i][g][0] for i in range(len(names))], 0))  # [rp_tot, k]
                    ub = torch.zeros(n_tot, rp_tot)
                    r0 = 0
                    for i, wi in enumerate(ws):
                        ub[r0:r0 + wi.shape[0], i * rp:(i + 1) * rp] = per_name[i][g][1]
                        r0 += wi.shape[0]
                    ups.append(ub)
                down, up = torch.cat(downs, 0), ups
            elif any(p is not None for p in per_name):
                raise NotImplementedError("partial LoRA coverage inside a fused projection")
        if perm is not None:
            w = w[perm]
            bias = bias[perm] if bias is not None else None
            if up is not None:
                up = [u[perm] for u in up]
        if up is not None and self.fuse_lora:
            # W_g = W + up_g down_g (controllora.py:728-737 `fuse_lora`, LoRA scale 1.0), one copy per weight group
            # stacked along N: a row segment of the batched pass then simply selects its copy (EsGemm.seg_b_noff).
            # Costs 2 extra copies of the Linear weights in HBM (~0.46 GB) and removes the 82 skinny down-projection
            # launches per step from the base pass.
            G = len(up)
            rp = rp_tot
            ws_g = [w] + [w + up[g] @ down[g * rp:(g + 1) * rp] for g in range(G)]
            bias_g = [bias] * (G + 1)
            if ln is not None:
                return self._fold_ln(ws_g, bias_g, ln, n_tot, block_n)
            bias_g = None if bias is None else torch.cat(bias_g)
            return Lin(self.mat(torch.cat(ws_g, 0)), None if bias_g is None else bias_g.to(self.device).contiguous(),
                       n_tot, None, None, 0, block_n, G + 1)
        if ln is not None:
            if up is not None:
                raise NotImplementedError("folded LayerNorm with an unfused LoRA update")
            return self._fold_ln([w], [bias], ln, n_tot, block_n)
        if up is not None:
            up = torch.cat(up, 0)
        return Lin(self.mat(w), None if bias is None else bias.to(self.device).contiguous(), n_tot,
                   None if down is None else self.mat(down), None if up is None else self.mat(up), rp_tot, block_n)

    def _fold_ln(self, ws_g, bias_g, ln, n_tot, block_n) -> Lin:
        """LayerNorm(x) W^T + b == rstd (x (W gamma)^T - mean colsum) + (b + W beta): one folded copy per weight group."""
        gamma, beta = [t.float().cpu() for t in ln]
        w_f, b_f, cs = [], [], []
        for wg, bg in zip(ws_g, bias_g):
            wf = (wg * gamma[None, :]).to(self.dtype)  # what the tensor cores multiply with
            w_f.append(wf)
            cs.append(wf.float().sum(1))
            b_f.append(wg @ beta + (bg if bg is not None else 0.0))
        return Lin(torch.cat(w_f, 0).to(self.device).contiguous(), torch.cat(b_f).to(self.device).contiguous(), n_tot,
                   None, None, 0, block_n, len(ws_g), torch.cat(cs).to(self.device).contiguous())

    def _conv_lora(self, name: str, k_cols: Optional[int] = None):
        """LoRAConv2dLayer of conv `name` per LoRA group: [(down [rp, taps * cin] tap-major like `conv3`, up [cout, rp])]
        with the rank zero-padded to a multiple of 8 (and the K columns to `k_cols`); None when the conv has no LoRA."""
        parts = []
        for sd in self.loras:
            dk, uk = f"{name}.lora_layer.down.weight", f"{name}.lora_layer.up.weight"
            if dk not in sd:
                return None
            d, u = sd[dk].float().cpu(), sd[uk].float().cpu()  # [r, cin, kh, kw], [cout, r, 1, 1]
            r = d.shape[0]
            rp = _pad8(r)
            d2 = d.permute(0, 2, 3, 1).reshape(r, -1)
            dp = torch.zeros(rp, k_cols or d2.shape[1])
            dp[:r, :d2.shape[1]] = d2
            up = torch.zeros(u.shape[0], rp)
            up[:, :r] = u.reshape(u.shape[0], r)
            parts.append((dp, up))
        return parts if parts else None

    def conv(self, name: str, k_cols: Optional[int] = None, n_pad: Optional[int] = None):
        """Conv weight `name` as a GEMM operand [cout, taps * cin] (tap-major) with its conv LoRA.  Returns
        (w, groups, lora): fused mode stacks [W; W + up_g down_g ...] along N (`_fuse_lora` of LoRACompatibleConv:
        W += (up.flatten(1) @ down.flatten(1)).reshape(W.shape)), unfused keeps (down, up, rp) for the K-extension."""
        w = self.sd[name + ".weight"].float().cpu()
        cout = w.shape[0]
        w2 = w.permute(0, 2, 3, 1).reshape(cout, -1)
        if k_cols is not None and k_cols != w2.shape[1]:
            wp = torch.zeros(cout, k_cols)
            wp[:, :w2.shape[1]] = w2
            w2 = wp
        parts = self._conv_lora(name, k_cols)
        if parts is None:
            return self.mat(w2), 1, None
        if self.fuse_lora:
            return self.mat(torch.cat([w2] + [w2 + u @ d for d, u in parts], 0)), 1 + len(parts), None
        rp = parts[0][0].shape[0]
        return self.mat(w2), 1, (self.mat(torch.cat([d for d, _ in parts], 0)), self.mat(torch.cat([u for _, u in parts], 0)), rp)

    def res(self, p: str) -> Res:
        sd = self.sd
        cout, cin = sd[f"{p}.conv1.weight"].shape[:2]
        w1, g1, l1 = self.conv(f"{p}.conv1")
        w2, g2, l2 = self.conv(f"{p}.conv2")
        groups = max(g1, g2)
        assert g1 == g2, "conv LoRA must cover conv1 and conv2 alike"
        lora = {}
        if l1 is not None:
            lora["conv1"], lora["conv2"] = l1, l2
        wsc = bsc = None
        b1 = self.f32(f"{p}.conv1.bias")
        b2 = self.f32(f"{p}.conv2.bias")
        if f"{p}.conv_shortcut.weight" in sd:
            wsc, gsc, lsc = self.conv(f"{p}.conv_shortcut")
            assert gsc == groups
            bsc = self.f32(f"{p}.conv_shortcut.bias")
            if lsc is not None:
                lora["conv_shortcut"] = lsc
            else:
                b2 = b2 + bsc  # the shortcut rides in conv2's accumulator (K-extension): one bias
        if groups > 1:  # the bias is indexed with the weight-row offset of the segment: one copy per group
            b1, b2 = b1.repeat(groups), b2.repeat(groups)
        return Res(self.f32(f"{p}.norm1.weight"), self.f32(f"{p}.norm1.bias"), w1, b1,
                   self.f32(f"{p}.norm2.weight"), self.f32(f"{p}.norm2.bias"), w2, b2, wsc, cin, cout,
                   groups=groups, lora=lora or None, bsc=bsc)

    def convw(self, name: str, k_cols: Optional[int] = None) -> ConvW:
        w, groups, lora = self.conv(name, k_cols)
        bias = self.f32(name + ".bias")
        cout = bias.shape[0]
        return ConvW(w, bias.repeat(groups) if groups > 1 else bias, cout, groups, lora)

    def tfm(self, p: str) -> Tfm:
        t = f"{p}.transformer_blocks.0"
        c = self.sd[f"{p}.norm.weight"].shape[0]
        n_ff = 8 * c
        bn = _geglu_block_n(n_ff)
        half = bn // 2
        idx = []
        for tile in range(n_ff // bn):
            idx += list(range(tile * half, (tile + 1) * half))
            idx += list(range(4 * c + tile * half, 4 * c + (tile + 1) * half))
        perm = torch.tensor(idx)
        ln = lambda n: (self.f32(f"{t}.{n}.weight"), self.f32(f"{t}.{n}.bias"))
        fold = (lambda n: (self.sd[f"{t}.{n}.weight"], self.sd[f"{t}.{n}.bias"])) if self.fold_ln else (lambda n: None)
        return Tfm(
            self.f32(f"{p}.norm.weight"), self.f32(f"{p}.norm.bias"),
            self.lin([f"{p}.proj_in"], [f"{p}.proj_in.bias"]),
            ln("norm1"),
            self.lin([f"{t}.attn1.to_q", f"{t}.attn1.to_k", f"{t}.attn1.to_v"], ln=fold("norm1")),
            self.lin([f"{t}.attn1.to_out.0"], [f"{t}.attn1.to_out.0.bias"]),
            ln("norm2"),
            self.lin([f"{t}.attn2.to_q"], ln=fold("norm2")),
            self.lin([f"{t}.attn2.to_k", f"{t}.attn2.to_v"]),
            self.lin([f"{t}.attn2.to_out.0"], [f"{t}.attn2.to_out.0.bias"]),
            ln("norm3"),
            self.lin([f"{t}.ff.net.0.proj"], [f"{t}.ff.net.0.proj.bias"], perm=perm, block_n=bn, ln=fold("norm3")),
            self.lin([f"{t}.ff.net.2"], [f"{t}.ff.net.2.bias"]),
            self.lin([f"{p}.proj_out"], [f"{p}.proj_out.bias"]),
            c,
            p,
        )


def merge_level_groups(nlev: int, n_up: int, stepping: bool, by_level: bool, split: int) -> List[List[int]]:
    """Residual levels (0 .. nlev - 2 = skips, nlev - 1 = mid) grouped into merge launches, in the order the decoder
    consumes them (mid, then skips nlev - 2 .. 0).  Outside the fused step (mode "residuals") one group; by_level: one
    group per decoder level -- (mid + the deepest level's n_up skips), then n_up skips each; else the levels >= split and
    the levels < split (split 0: one group)."""
    order = list(reversed(range(nlev)))
    if not stepping:
        return [order]
    if by_level:
        rest = order[1 + n_up:]
        return [order[:1 + n_up]] + [rest[i:i + n_up] for i in range(0, len(rest), n_up)]
    split = max(0, min(split, nlev))
    return [g for g in (list(reversed(range(split, nlev))), list(reversed(range(split)))) if g]


@dataclass(frozen=True)
class StepGeometry:
    base_nets: Tuple[Optional[int], ...]  # net index per image block of the base pass (None = the UNet's own rows)
    pose_nets: Tuple[int, ...]            # net index per image block of the pose pass
    seg: Tuple[int, int, int]             # images of (UNet | agnostic | clothes) rows in the base pass
    btag: str
    ptag: str


@dataclass
class EncoderW:
    conv_in: ConvW  # w [groups * c0, 64] (im2col of 4 channels x 9 taps, zero padded)
    down_res: List[List[Res]]
    down_tfm: List[List[Optional[Tfm]]]
    down_conv: List[Optional[ConvW]]
    mid_res: List[Res]
    mid_tfm: Tfm
    # time path: per weight-set group (LoRA folded into these tiny-M linears at pack time)
    te1: List[Tuple[torch.Tensor, torch.Tensor]]
    te2: List[Tuple[torch.Tensor, torch.Tensor]]
    temb_w: List[torch.Tensor]  # per group [sum cout, temb_dim]
    temb_b: List[torch.Tensor]
    temb_cols: int = 0


class DenoiseEngine:
    """One B200, one weight replica, a fixed (rows, h, w) problem."""

    def __init__(self, cfg, unet_sd, lora_sds, pose_sd, merge_sd, *, rows: int, h: int, w: int,
                 dtype=torch.float16, device="cuda", n_text: int = 77, pattern=SUPPORTED_PATTERN,
                 use_graph: bool = False, fuse_lora: Optional[bool] = None):
        if tuple(pattern) != SUPPORTED_PATTERN:
            raise NotImplementedError(f"load_pattern {pattern}: only {SUPPORTED_PATTERN} (app.py:40) is supported")
        if not torch.cuda.is_available():
            raise RuntimeError("DenoiseEngine needs a CUDA device: there is no CPU fallback")
        import os

        from .ext import load

        lib = load()  # fail loudly if the native library is missing
        # programmatic dependent launch between consecutive kernels of the step (ES_PDL=0 disables)
        lib.es_set_pdl(1 if os.environ.get("ES_PDL", "1") != "0" else 0)
        self.cfg = cfg = C.UNetConfig.from_any(cfg)
        self.B, self.h, self.w, self.dtype, self.dev, self.n_text = rows, h, w, dtype, torch.device(device), n_text
        self.levels = C.level_sizes(h, w, len(cfg.block_out_channels))
        self.use_graph = use_graph
        self._bufs: Dict[str, torch.Tensor] = {}
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._graph_launches: Dict[tuple, int] = {}
        self.fuse_gn_stats = os.environ.get("ES_FUSE_GN", "1") != "0"
        # ControlLoRA update: "fused" = one fused weight copy per LoRA group (default), "unfused" = rank-r update as
        # extra K-blocks of the accumulator (t = x down^T, then [x | t] [W | up]^T)
        self.fuse_lora = os.environ.get("ES_LORA", "fused") != "unfused" if fuse_lora is None else bool(fuse_lora)
        # zero-convs + merge on the side stream: 0 = after both encoders (in the order the decoder consumes them),
        # 1 = level by level as the encoders produce them, 2 = hybrid: the small levels (32x32 and below) as they are
        # produced, the three heavy 64x64-level merges afterwards, under the decoder's latency-bound deep levels
        # BasicTransformerBlock LayerNorms folded around the GEMMs (EsGemm.rowstat_out / ln_rowstat): the producing
        # projection accumulates per-row (sum, sumsq), the consuming one normalises in its epilogue -- no LayerNorm
        # kernel and no normalised copy of the hidden state in memory.  Needs whole 128-row tiles per row segment.
        self.fold_ln = (os.environ.get("ES_FOLD_LN", "1") != "0" and self.fuse_lora
                        and (rows * self.levels[-1][0] * self.levels[-1][1]) % 128 == 0)
        # merge levels >= merge_split form the first group (one launch per phase, deep levels: the decoder starts on
        # them), levels < merge_split the second (the heavy 64x64-level merges, under the decoder's deep levels);
        # 0 = a single group
        self.merge_split = int(os.environ.get("ES_MERGE_SPLIT", "3"))
        # next-layer weight prefetch into L2 (ops.WeightPrefetchPlan): "0" off, "1" every GEMM, "decoder" the UNet decoder only
        self.weight_prefetch = os.environ.get("ES_WEIGHT_PREFETCH", "0")
        # ES_MERGE_BY_LEVEL=1 (default): one merge group per decoder level instead (4 groups = 12 launches at SD1.5's
        # 13 residual levels): the decoder starts after the 8x8-level group alone -- 165 us after the encoders instead of
        # 350 us (tools/timeline.py), 9.81 vs 9.92 ms/step
        self.merge_by_level = os.environ.get("ES_MERGE_BY_LEVEL", "1") != "0"
        # ES_ZC_CHAIN=1: after its encoder pass each chain runs its own zero-convs (deepest group first) on a stream of
        # its own instead of queueing them on the merge stream in front of each group's merge.  Default off -- measured
        # 9.91 vs 9.81 ms/step: the extra concurrency only competes with the decoder's first kernels
        # "first": only the first (deepest) group's zero-convs go to the chains' streams -- the decoder waits for them
        self.zc_chain = os.environ.get("ES_ZC_CHAIN", "0")
        self.merge_z16 = os.environ.get("ES_MERGE_Z16", "1") != "0"  # fp16 engines keep the merge's z tensor in fp16
        self._stats_of = {}
        self._cat_slot = {}
        self._kv_recompute = False  # True: redo the text K/V projections inside every step (reference behaviour)
        self._temb_ready = None     # event of the base pass's time path while it is pending on the side stream
        self.launches_per_step = 0
        for i, c in enumerate(cfg.block_out_channels):
            if c % 8 or c % cfg.norm_num_groups or (c // cfg.num_heads) % 8:
                raise ValueError(f"channel count {c} unsupported (needs %8, %groups, head_dim %8)")
        ops.set_gemm_workspace(256 << 20, self.dev)  # split-K scratch of the main stream
        # stream priorities (ES_PRIO = "<chains>,<merge>", lower number = higher priority, 0 = default)
        prio = [int(v) for v in os.environ.get("ES_PRIO", "0,0").split(",")]
        self._chain_streams = [torch.cuda.Stream(device=self.dev, priority=prio[0])]  # the pose pass
        self._merge_stream = torch.cuda.Stream(device=self.dev, priority=prio[1])
        # ES_ZC_EARLY=1: zero-convs run level by level on their own streams as the two encoder passes produce their
        # residual levels.  Default off: after both passes, on the merge stream -- measured 10.09 vs 10.17 ms/step (the
        # 26 small GEMMs compete with the encoders' kernels for SMs while those are on the critical path).
        self.zc_early = os.environ.get("ES_ZC_EARLY", "0") != "0"
        self._zc_streams = [torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)]
        self._side_ws = []
        for st in self._chain_streams + [self._merge_stream] + self._zc_streams:  # concurrent launches must not share split-K scratch
            ws = torch.zeros(128 << 20, dtype=torch.uint8, device=self.dev)
            self._side_ws.append(ws)
            ops.set_stream_workspace(st, ws)
        self._pack(unet_sd, lora_sds, pose_sd, merge_sd)
        self._alloc_static()
        # per-shape GEMM tile / split-K selection: committed table first, measure whatever is missing during the
        # first (eager) step on this device
        self._tuned_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuned_b200.json")
        if os.environ.get("ES_AUTOTUNE", "1") != "0":
            if os.environ.get("ES_RETUNE", "0") == "0":
                ops.TUNER.load(self._tuned_path)
            ops.TUNER.enabled = True
        self._tune = os.environ.get("ES_AUTOTUNE", "1") != "0"
        self._tuning_done = False
        self._kv_masks = set()      # active-net masks whose text K/V projections exist for the current prompt
        self._scale_host = None

    # ------------------------------------------------------------------------------------ packing
    def _pack_encoder(self, sd, loras) -> EncoderW:
        cfg = self.cfg
        P = _Packer(sd, loras, self.dtype, self.dev, self.fuse_lora, self.fold_ln)
        c0 = cfg.block_out_channels[0]
        down_res, down_tfm, down_conv = [], [], []
        for i in range(len(cfg.block_out_channels)):
            down_res.append([P.res(f"down_blocks.{i}.resnets.{j}") for j in range(cfg.layers_per_block)])
            down_tfm.append([P.tfm(f"down_blocks.{i}.attentions.{j}") if cfg.down_has_attn[i] else None
                             for j in range(cfg.layers_per_block)])
            k = f"down_blocks.{i}.downsamplers.0.conv"
            down_conv.append(P.convw(k) if k + ".weight" in sd else None)
        mid_res = [P.res("mid_block.resnets.0"), P.res("mid_block.resnets.1")]
        mid_tfm = P.tfm("mid_block.attentions.0")
        all_res = [r for lvl in down_res for r in lvl] + mid_res
        return self._finish_time_path(EncoderW(P.convw("conv_in", 64), down_res, down_tfm, down_conv,
                                               mid_res, mid_tfm, [], [], [], []), sd, loras, all_res,
                                      [f"down_blocks.{i}.resnets.{j}" for i in range(len(cfg.block_out_channels))
                                       for j in range(cfg.layers_per_block)] + ["mid_block.resnets.0",
                                                                                "mid_block.resnets.1"])

    def _finish_time_path(self, E: EncoderW, sd, loras, all_res: List[Res], res_names: List[str]) -> EncoderW:
        """Tiny-M linears (time_embedding, time_emb_proj): one weight copy per group with LoRA folded in
        (W + up @ down, exactly `_fuse_lora`, diffusers models/lora.py) -- M = rows, so these are GEMVs."""
        off = 0
        for r in all_res:
            r.temb_off = off
            off += r.cout
        E.temb_cols = off

        def fused(name, lsd):
            w = sd[name + ".weight"].float().cpu()
            if lsd is not None and f"{name}.lora_layer.down.weight" in lsd:
                w = w + lsd[f"{name}.lora_layer.up.weight"].float().cpu() @ lsd[f"{name}.lora_layer.down.weight"].float().cpu()
            return w

        groups = [None] + list(loras)
        for lsd in groups:
            E.te1.append((self._mat(fused("time_embedding.linear_1", lsd)), self._f32(sd["time_embedding.linear_1.bias"])))
            E.te2.append((self._mat(fused("time_embedding.linear_2", lsd)), self._f32(sd["time_embedding.linear_2.bias"])))
            E.temb_w.append(self._mat(torch.cat([fused(n + ".time_emb_proj", lsd) for n in res_names], 0)))
            E.temb_b.append(self._f32(torch.cat([sd[n + ".time_emb_proj.bias"].float().cpu() for n in res_names], 0)))
        return E

    def _mat(self, t):
        return t.detach().to("cpu", self.dtype).contiguous().to(self.dev)

    def _pack_pose_embedder(self, sd):
        """ControlNetConditioningEmbedding of the openpose net (diffusers controlnet.py; reached from
        CachedControlNetModel.preprocess_image, controllora.py:289-290, and :199-201 for raw images): conv_in, pairs
        of (3x3, 3x3 stride 2) blocks, conv_out -- [(weight [cout, 9 * cin_pad], bias, cin_pad, cout, stride, silu)].
        None when the checkpoint carries no embedder."""
        p = "controlnet_cond_embedding"
        if f"{p}.conv_in.weight" not in sd:
            return None
        names = [(f"{p}.conv_in", 1, True)]
        i = 0
        while f"{p}.blocks.{i}.weight" in sd:
            names.append((f"{p}.blocks.{i}", 2 if i % 2 else 1, True))
            i += 1
        names.append((f"{p}.conv_out", 1, False))
        layers = []
        for name, stride, silu in names:
            w = sd[name + ".weight"].float().cpu()  # [cout, cin, 3, 3]
            cout, cin = w.shape[:2]
            cin_pad = _pad8(cin)
            wp = torch.zeros(cout, 3, 3, cin_pad)
            wp[..., :cin] = w.permute(0, 2, 3, 1)
            layers.append((self._mat(wp.reshape(cout, 9 * cin_pad)), self._f32(sd[name + ".bias"]), cin_pad, cout, stride,
                           silu))
        return layers

    def _f32(self, t):
        return t.detach().to("cpu", torch.float32).contiguous().to(self.dev)

    def _pack(self, unet_sd, lora_sds, pose_sd, merge_sd):
        """Any component may be None (a stand-alone ControlNet only needs its own encoder and zero-convs): `unet_sd`
        (base encoder + decoder), `pose_sd` (plain ControlNet encoder), `merge_sd` (EdgeStyle merge blocks); `lora_sds`
        holds one state dict per ControlLoRA weight group over the UNet's weights."""
        cfg = self.cfg
        self.full = unet_sd is not None and pose_sd is not None and merge_sd is not None and len(lora_sds) == 2
        self.enc_base = self._pack_encoder(unet_sd, lora_sds) if unet_sd is not None else None
        self.enc_pose = self._pack_encoder(pose_sd, []) if pose_sd is not None else None
        self.pose_embed = self._pack_pose_embedder(pose_sd) if pose_sd is not None else None
        self.res_shapes = C.residual_shapes(cfg, self.h, self.w)
        self.up_res, self.up_tfm, self.up_conv = [], [], []
        self.dec_temb_cols = 0
        if self.full:  # UNet decoder
            P = _Packer(unet_sd, [], self.dtype, self.dev, True, self.fold_ln)
            nb = len(cfg.block_out_channels)
            rev_attn = list(reversed(cfg.down_has_attn))
            names = []
            for i in range(nb):
                self.up_res.append([P.res(f"up_blocks.{i}.resnets.{j}") for j in range(cfg.layers_per_block + 1)])
                names += [f"up_blocks.{i}.resnets.{j}" for j in range(cfg.layers_per_block + 1)]
                self.up_tfm.append([P.tfm(f"up_blocks.{i}.attentions.{j}") if rev_attn[i] else None
                                    for j in range(cfg.layers_per_block + 1)])
                k = f"up_blocks.{i}.upsamplers.0.conv"
                self.up_conv.append((P.conv3(k + ".weight"), P.f32(k + ".bias")) if k + ".weight" in unet_sd else None)
            dec_res = [r for lvl in self.up_res for r in lvl]
            off = self.enc_base.temb_cols
            for r in dec_res:
                r.temb_off = off
                off += r.cout
            self.dec_temb_cols = off - self.enc_base.temb_cols
            # decoder time_emb_proj appended to the UNet group's (group 0) concatenated matrix
            self.enc_base.temb_w[0] = torch.cat(
                [self.enc_base.temb_w[0]] + [self._mat(unet_sd[n + ".time_emb_proj.weight"].float()) for n in names], 0)
            self.enc_base.temb_b[0] = torch.cat(
                [self.enc_base.temb_b[0]] + [self._f32(unet_sd[n + ".time_emb_proj.bias"].float()) for n in names], 0)
            self.norm_out = (P.f32("conv_norm_out.weight"), P.f32("conv_norm_out.bias"))
            co = cfg.out_channels
            wco = torch.zeros(16, 9 * cfg.block_out_channels[0])
            wco[:co] = unet_sd["conv_out.weight"].float().cpu().permute(0, 2, 3, 1).reshape(co, -1)
            self.conv_out = (self._mat(wco), self._f32(torch.cat([unet_sd["conv_out.bias"].float().cpu(),
                                                                 torch.zeros(16 - co)])))
        # zero convs: base pass stacks the LoRA groups ([agn; clo]) along N, pose separate
        zc = C.zero_conv_channels(cfg) + [cfg.block_out_channels[-1]]
        keys = [f"controlnet_down_blocks.{i}" for i in range(len(zc) - 1)] + ["controlnet_mid_block"]
        self.zero_base, self.zero_pose = [], []
        for k, c in zip(keys, zc):
            if lora_sds:
                wb = torch.cat([l[k + ".weight"].float().cpu().reshape(c, c) for l in lora_sds], 0)
                bb = torch.cat([l[k + ".bias"].float().cpu() for l in lora_sds], 0)
                self.zero_base.append((self._mat(wb), self._f32(bb)))
            if pose_sd is not None:
                self.zero_pose.append((self._mat(pose_sd[k + ".weight"].float().reshape(c, c)), self._f32(pose_sd[k + ".bias"])))
        # merge blocks
        self.merge = []
        if merge_sd is not None:
            pfx = [f"multi_controlnet_down_blocks.{i}." for i in range(len(self.res_shapes) - 1)] + ["multi_controlnet_mid_block."]
            for p, (c, hh, ww) in zip(pfx, self.res_shapes):
                sub = {k[len(p):]: v for k, v in merge_sd.items() if k.startswith(p)}
                self.merge.append(pack_merge_block(sub, c, hh, ww, self.dtype, self.dev))

    # ------------------------------------------------------------------------------------ buffers
    def buf(self, name: str, rows: int, cols: int, dtype=None) -> torch.Tensor:
        dtype = dtype or self.dtype
        key = f"{name}:{rows}x{cols}:{dtype}"
        t = self._bufs.get(key)
        if t is None:
            t = torch.zeros(rows, cols, device=self.dev, dtype=dtype)
            self._bufs[key] = t
        return t

    def _alloc_static(self):
        cfg, B = self.cfg, self.B
        hw = self.h * self.w
        c0 = cfg.block_out_channels[0]
        self.sample_in = torch.zeros(B, cfg.in_channels, self.h, self.w, device=self.dev, dtype=torch.float32)
        self.t_in = torch.zeros(B, device=self.dev, dtype=torch.float32)
        self.eps_out = torch.zeros(B, cfg.out_channels, self.h, self.w, device=self.dev, dtype=torch.float32)
        self.conds = self.buf("conds", 6 * B * hw, c0)          # nets 0..5, each [B*hw, c0]
        self.ctx_base = self.buf("ctx_base", 4 * B * self.n_text, cfg.cross_attention_dim)
        self.ctx_pose = self.ctx_base[: 3 * B * self.n_text]
        self.ctx_dec = self.ctx_base[: B * self.n_text]
        # one pool of reduction scratch for the whole step, zeroed by ONE memset at the start of the step: every
        # GroupNorm call / merge block takes its own slot (no per-call memset launches inside the graph)
        self._gn_slot_floats = 4 * B * cfg.norm_num_groups * 2
        gn_bytes = 256 * self._gn_slot_floats * 4
        merge_bytes = 32 * B * 4 * 8
        # folded LayerNorm: [rows, 2] fp32 (sum, sumsq) per LayerNorm instance of the step
        ln_rows = 0
        if self.fold_ln:
            n_lvl = len(cfg.block_out_channels)
            for i, (H, W) in enumerate(self.levels):
                if cfg.down_has_attn[i]:
                    ln_rows += 3 * H * W * (7 * B * cfg.layers_per_block + B * (cfg.layers_per_block + 1))
            H, W = self.levels[-1]
            ln_rows += 3 * H * W * 7 * B  # mid blocks of the two encoder passes
        self.scratch = torch.zeros(gn_bytes + merge_bytes + ln_rows * 8, device=self.dev, dtype=torch.uint8)
        self.gn_pool = self.scratch[:gn_bytes].view(torch.float32)
        self.merge_pool = self.scratch[gn_bytes:gn_bytes + merge_bytes].view(torch.float64).view(32, B, 4)
        self.ln_pool = self.scratch[gn_bytes + merge_bytes:].view(torch.float32)
        self._gn_next = 0
        self._merge_next = 0
        self._ln_next = 0
        self.coef = torch.zeros(4, device=self.dev, dtype=torch.float32)
        self.cond_scale_dev = torch.ones(6, device=self.dev, dtype=torch.float32)  # read by the merge kernels
        self.guidance = torch.ones(max(B // 2, 1), device=self.dev, dtype=torch.float32)

    # ------------------------------------------------------------------------------------ inputs
    def set_prompt(self, prompt_embeds: torch.Tensor):
        """prompt_embeds [B, n_text, ctx] (negative rows first, edgestyle_pipeline.py:330)."""
        B, nt = self.B, self.n_text
        assert prompt_embeds.shape == (B, nt, self.cfg.cross_attention_dim), prompt_embeds.shape
        if not self.full:
            raise RuntimeError("set_prompt belongs to the fused step; a stand-alone net takes the prompt per call")
        pe = prompt_embeds.to(device=self.dev, dtype=self.dtype).reshape(B * nt, -1)
        for g in range(4):
            self.ctx_base[g * B * nt:(g + 1) * B * nt].copy_(pe)
        masks = self._kv_masks or {(True,) * 6}
        self._kv_masks = set()
        for T in (t for lvl in self.up_tfm for t in lvl if t is not None):
            kv = self.buf(f"kv.dec.{T.uid}", B * nt, 2 * T.c)
            self._lin(T.kv2, self.ctx_dec, kv, nt, None, "dec.kv")
        for m in sorted(masks, reverse=True):
            self._ensure_text_kv(m)

    def _geometry(self, active: Sequence[bool]) -> "StepGeometry":
        """Image blocks of the two batched encoder passes for the set of contributing nets: the base (UNet weights)
        pass carries the UNet rows plus one block per active ControlLoRA net (agnostic = net 0, clothes = nets 2, 4),
        the pose pass one block per active openpose net (1, 3, 5)."""
        B = self.B
        base = [None] + [k for k in (0, 2, 4) if active[k]]
        pose = [k for k in (1, 3, 5) if active[k]]
        seg = (B, B if active[0] else 0, B * (int(active[2]) + int(active[4])))
        m = "".join("1" if a else "0" for a in active)
        return StepGeometry(tuple(base), tuple(pose), seg, "b" + m[0::2], "p" + m[1::2])

    def _ensure_text_kv(self, active):
        """attn2.to_k / to_v of every encoder transformer on the (step-invariant) prompt embeddings, LoRA included, for
        the image blocks of this set of active nets.  Computed once per (prompt, set): never inside a captured step."""
        active = tuple(bool(a) for a in active)
        if active in self._kv_masks:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("text K/V projections must exist before a step is captured")
        B, nt = self.B, self.n_text
        geo = self._geometry(active)
        for tag, E, nb, seg in ((geo.btag, self.enc_base, len(geo.base_nets), geo.seg),
                                (geo.ptag, self.enc_pose, len(geo.pose_nets), None)):
            if nb == 0:
                continue
            ctx = (self.ctx_base if E is self.enc_base else self.ctx_pose)[: nb * B * nt]
            tf = [t for lvl in E.down_tfm for t in lvl if t is not None] + [E.mid_tfm]
            for T in tf:
                kv = self.buf(f"kv.{tag}.{T.uid}", nb * B * nt, 2 * T.c)
                self._lin(T.kv2, ctx, kv, nt, seg, tag + ".kv")
        self._kv_masks.add(active)

    def _convw_block(self, Cw: ConvW, col, out, grp: int, tag: str, residual=None):
        """conv_in of ONE image block (flat GEMM on the im2col of the sample) with the weights of LoRA group `grp`
        (0 = plain UNet weights, 1 = agnostic, 2 = clothes): fused copy, or the rank-r K-extension."""
        co = Cw.cout
        if Cw.groups > 1:
            return ops.gemm(col, Cw.w, co, out=out, bias=Cw.bias, residual=residual,
                            segs=([0, col.shape[0]], [grp * co], None))
        if Cw.lora is not None and grp > 0:
            down, up, rp = Cw.lora
            t = self.buf(f"{tag}.lora_t", col.shape[0], rp)
            ops.gemm(col, down, rp, out=t, segs=([0, col.shape[0]], [(grp - 1) * rp], None))
            return ops.gemm(col, Cw.w, co, out=out, bias=Cw.bias, residual=residual, a2=t, b2=up,
                            segs=([0, col.shape[0]], [0], [(grp - 1) * co]))
        return ops.gemm(col, Cw.w, co, out=out, bias=Cw.bias, residual=residual)

    def set_conditioning(self, conds: Sequence[torch.Tensor]):
        """Six cached conditioning embeddings [B, c0, h, w] (prepare_image, edgestyle_pipeline.py:629-664)."""
        B, hw = self.B, self.h * self.w
        assert len(conds) == 6
        for k, c in enumerate(conds):
            assert c.shape == (B, self.cfg.block_out_channels[0], self.h, self.w), c.shape
            ops.nchw_to_nhwc(c.to(device=self.dev, dtype=torch.float32).contiguous(), self.conds[k * B * hw:(k + 1) * B * hw])

    def _begin_step_scratch(self):
        self._stats_of = {}
        self._cat_slot = {}
        self.scratch.zero_()
        self._gn_next = 0
        self._merge_next = 0
        self._ln_next = 0

    def _ln_slot(self, rows: int):
        """[rows, 2] fp32 (sum, sumsq), zero at the start of the step (one memset for the whole pool)."""
        n = 2 * rows
        if self._ln_next + n > self.ln_pool.numel():
            raise RuntimeError("LayerNorm statistics pool exhausted")
        t = self.ln_pool[self._ln_next:self._ln_next + n]
        self._ln_next += n
        return t

    def _merge_slot(self):
        s = self.merge_pool[self._merge_next]
        self._merge_next += 1
        return s

    # ------------------------------------------------------------------------------------ layers
    def _gn_slot(self, imgs):
        G = self.cfg.norm_num_groups
        if self._gn_next >= 256:
            raise RuntimeError("GroupNorm scratch pool exhausted")
        ws = self.gn_pool[self._gn_next * self._gn_slot_floats:][: imgs * G * 2].view(imgs, G, 2)
        self._gn_next += 1
        return ws

    def _stats_for(self, out, imgs, hw):
        """Slot in which the GEMM producing `out` accumulates GroupNorm statistics (None when not fusable: a strided
        view of a concat buffer, or a level whose images are not a multiple of 32 rows)."""
        if not self.fuse_gn_stats or hw % 32 != 0 or out.stride(0) != out.shape[1]:
            return None
        ws = self._gn_slot(imgs)
        self._stats_of[(out.data_ptr(), out.shape[0], out.shape[1])] = ws
        return ws

    def _gn_kw(self, out, imgs, hw) -> dict:
        """GroupNorm-statistics arguments for the GEMM that produces `out`: a dense tensor gets its own slot; the
        x half of a registered decoder concat buffer accumulates into the concat's slot (EsGemm.gn_cpg / gn_col0)."""
        G = self.cfg.norm_num_groups
        if out.stride(0) != out.shape[1]:
            reg = self._cat_slot.get(out.data_ptr())
            if reg is None:
                return {}
            ws, cpg = reg
            return dict(gn_ws=ws, gn_groups=G, gn_cpg=cpg, gn_col0=0)
        return dict(gn_ws=self._stats_for(out, imgs, hw), gn_groups=G)

    def _register_cat(self, cbuf, imgs, hw) -> bool:
        """Decoder [x | skip] concat buffer: both halves are written by kernels that can accumulate the GroupNorm
        statistics of the first resnet norm (GEMM epilogue for x, merge phase 3 for skip) -> no statistics pass."""
        G = self.cfg.norm_num_groups
        C = cbuf.shape[1]
        if not self.fuse_gn_stats or hw % 32 != 0 or C % G or (C // G) < 8 or (C // G) % 2:
            return False
        ws = self._gn_slot(imgs)
        self._stats_of[(cbuf.data_ptr(), cbuf.shape[0], C)] = ws
        self._cat_slot[cbuf.data_ptr()] = (ws, C // G)
        return True

    def _gn(self, x, out, g, b, imgs, hw, silu, eps=None):
        G = self.cfg.norm_num_groups
        eps = self.cfg.norm_eps if eps is None else eps
        ws = self._stats_of.pop((x.data_ptr(), x.shape[0], x.shape[1]), None)
        if ws is not None:
            ops.groupnorm(x, out, g, b, ws, imgs, hw, G, eps, silu, stats_ready=True)
        else:
            ops.groupnorm(x, out, g, b, self._gn_slot(imgs), imgs, hw, G, eps, silu, zero_ws=False)
        return out

    def _lin(self, L: Lin, a, out, rows_per_img: int, lora_seg_imgs: Optional[Sequence[int]], tag: str, ln_stat=None,
             **ep):
        """out = a @ L.w^T (+ LoRA per row segment) with the fused epilogue `ep`; ln_stat: row statistics of `a` when
        the LayerNorm in front of this projection is folded into it (L.colsum)."""
        if ep.get("gn_ws") is not None:
            ep["rows_per_img"] = rows_per_img  # fused GroupNorm statistics are per image
        if L.colsum is not None:
            assert ln_stat is not None, "folded-LayerNorm projection called without row statistics"
            ep["ln"] = (ln_stat, L.colsum, a.shape[1], 1e-5)
        if L.groups > 1:
            M = a.shape[0]
            if lora_seg_imgs is None:
                segs = ([0, M], [0], None)
            else:
                n0, n1, n2 = [s * rows_per_img for s in lora_seg_imgs]  # rows of: no-LoRA | group 0 | group 1
                segs = ([0, n0, n0 + n1, n0 + n1 + n2], [0, L.n, 2 * L.n], None)
            ops.gemm(a, L.w, L.n, out=out, bias=L.bias, block_n=L.block_n, segs=segs, **ep)
        elif L.down is not None and lora_seg_imgs is not None and (lora_seg_imgs[1] + lora_seg_imgs[2]) > 0:
            n0, n1, n2 = [s * rows_per_img for s in lora_seg_imgs]  # rows of: no-LoRA | group 0 | group 1
            M = a.shape[0]
            t = self.buf(f"{tag}.lora_t", M, L.rp)
            ops.gemm(a[n0:], L.down, L.rp, out=t[n0:], segs=([0, n1, n1 + n2], [0, L.rp], None))
            ops.gemm(a, L.w, L.n, out=out, bias=L.bias, a2=t, b2=L.up, block_n=L.block_n,
                     segs=([0, n0, n0 + n1, n0 + n1 + n2], [0, 0, 0], [-1, 0, L.n]), **ep)
        else:
            ops.gemm(a, L.w, L.n, out=out, bias=L.bias, block_n=L.block_n, **ep)
        return out

    @staticmethod
    def _imgs_per_tile(H: int, W: int) -> int:
        """Images one 128-row conv tile of es_gemm covers (gemm.cu: bw x bh x bn boxes): > 1 only on tiny levels."""
        bw = W if W < 128 else 128
        bh = max(128 // bw, 1)
        if bh <= H:
            return 1
        hh = 1
        while hh < H:
            hh <<= 1
        return max(bh // hh, 1)

    def _conv_seg(self, a, w, n, *, out, H, W, imgs, taps, seg, noff, noff2=None, a2=None, b2=None, rowvec=None,
                  residual=None, gn_ws=None, **kw):
        """One convolution (taps 9: implicit GEMM over [imgs, H, W]; taps 1: flat) whose weight rows depend on the
        image segment: seg = image counts of (UNet | agnostic | clothes) rows, noff / noff2 = weight-row offset of
        each segment in `w` / `b2` (-1 in noff2: no second source for that segment).  One launch when the segment
        boundaries fall on tile boundaries, else one launch per segment over its image range."""
        hw = H * W
        if seg is None:
            seg, noff, noff2 = (imgs, 0, 0), (noff[0], 0, 0), None if noff2 is None else (noff2[0], 0, 0)
        starts = [0, seg[0], seg[0] + seg[1], seg[0] + seg[1] + seg[2]]
        assert starts[-1] == imgs, (seg, imgs)
        common = dict(taps=taps, a2=a2, b2=b2, rowvec=rowvec, residual=residual, gn_ws=gn_ws, **kw)
        unit = 1 if taps == 9 else hw  # segment starts: images (implicit conv) or rows (flat)
        per_tile = self._imgs_per_tile(H, W) if taps == 9 else 1  # flat GEMMs clip segment tails themselves
        if all(b % per_tile == 0 for b in starts[1:3]):
            segs = ([b * unit for b in starts], list(noff), None if noff2 is None else list(noff2))
            if taps == 9:
                common["whn"] = (W, H, imgs)
            elif gn_ws is not None or rowvec is not None:
                common["rows_per_img"] = hw
            return ops.gemm(a, w, n, out=out, segs=segs, **common)
        for i in range(3):  # ragged segments (e.g. one image per segment on a level whose tile holds two)
            i0, i1 = starts[i], starts[i + 1]
            if i1 == i0:
                continue
            sub = dict(common)
            for key in ("a2", "residual"):
                if sub[key] is not None:
                    sub[key] = sub[key][i0 * hw:i1 * hw]
            for key in ("rowvec", "gn_ws"):
                if sub[key] is not None:
                    sub[key] = sub[key][i0:i1]
            if noff2 is not None and noff2[i] < 0:
                sub["a2"] = sub["b2"] = None
            if taps == 9:
                sub["whn"] = (W, H, i1 - i0)
            elif gn_ws is not None or rowvec is not None:
                sub["rows_per_img"] = hw
            ops.gemm(a[i0 * hw:i1 * hw], w, n, out=out[i0 * hw:i1 * hw],
                     segs=([0, (i1 - i0) * unit], [noff[i]], None if sub["a2"] is None else [max(noff2[i], 0) if noff2 else 0]),
                     **sub)
        return out

    def _lora_t(self, x, lora, seg, H, W, imgs, taps, tag, c1):
        """Unfused conv LoRA: t = conv_down_g(x) over the LoRA rows only ([imgs * hw, rp]; the UNet rows are skipped)."""
        down, up, rp = lora
        hw = H * W
        n0 = seg[0]
        t = self.buf(f"{tag}.lora_t", imgs * hw, rp)
        if imgs - n0 > 0:
            self._conv_seg(x[n0 * hw:], down, rp, out=t[n0 * hw:], H=H, W=W, imgs=imgs - n0, taps=taps,
                           seg=(0, seg[1], seg[2]), noff=(0, 0, rp), c1=c1)
        return t

    def _convw(self, Cw: ConvW, col, out, H, W, imgs, seg, tag, **kw):
        """A stand-alone convolution run as a flat GEMM on an im2col matrix (conv_in, stride-2 down-sampler) with its
        conv LoRA: fused weight copy per image segment, or the K-extension t = col @ down^T, [col | t] [W | up]^T."""
        co = Cw.cout
        if seg is None or (Cw.groups == 1 and Cw.lora is None):
            if kw.get("gn_ws") is not None:
                kw["rows_per_img"] = H * W
            return ops.gemm(col, Cw.w, co, out=out, bias=Cw.bias, **kw)
        if Cw.groups > 1:
            return self._conv_seg(col, Cw.w, co, out=out, H=H, W=W, imgs=imgs, taps=1, seg=seg, noff=(0, co, 2 * co),
                                  bias=Cw.bias, **kw)
        t = self._lora_t(col, Cw.lora, seg, H, W, imgs, 1, tag, col.shape[1])
        return self._conv_seg(col, Cw.w, co, out=out, H=H, W=W, imgs=imgs, taps=1, seg=seg, noff=(0, 0, 0),
                              noff2=(-1, 0, co), bias=Cw.bias, a2=t, b2=Cw.lora[1], **kw)

    def _resnet(self, R: Res, x, imgs, H, W, temb, out, tag, seg=None):
        """ResnetBlock2D.  seg = image counts of (UNet | agnostic | clothes) rows when the convolutions carry a conv
        LoRA (lora_conv2d_rank > 0, /root/reference/model/controllora.py:561-575): fused weight copies are selected
        per image segment, the unfused update rides as a K-extension (source 2 = down-conv output, B2 = up)."""
        M = imgs * H * W
        g1 = self.buf(f"{tag}.gn1", M, R.cin)
        self._gn(x, g1, R.n1g, R.n1b, imgs, H * W, True)
        hbuf = self.buf(f"{tag}.h", M, R.cout)
        G = self.cfg.norm_num_groups
        if self._temb_ready is not None and tag.startswith("base"):
            st = torch.cuda.current_stream()
            if st.cuda_stream not in self._temb_waited:  # time path of the base pass (enqueued on the side stream)
                st.wait_event(self._temb_ready)
                self._temb_waited.add(st.cuda_stream)
        rowvec = temb[:, R.temb_off:R.temb_off + R.cout]
        co = R.cout
        if (R.groups == 1 and R.lora is None) or seg is None:
            ops.gemm(g1, R.w1, co, out=hbuf, taps=9, whn=(W, H, imgs), bias=R.b1, rowvec=rowvec, c1=R.cin,
                     gn_ws=self._stats_for(hbuf, imgs, H * W), gn_groups=G)
            g2 = self.buf(f"{tag}.gn2", M, co)
            self._gn(hbuf, g2, R.n2g, R.n2b, imgs, H * W, True)
            okw = self._gn_kw(out, imgs, H * W)
            if R.wsc is not None:
                ops.gemm(g2, R.w2, co, out=out, taps=9, whn=(W, H, imgs), bias=R.b2, a2=x, b2=R.wsc, c1=co, **okw)
            else:
                ops.gemm(g2, R.w2, co, out=out, taps=9, whn=(W, H, imgs), bias=R.b2, residual=x, c1=co, **okw)
            return out
        if R.groups > 1:  # fused copies [W; W + up_1 down_1; W + up_2 down_2] stacked along N
            noff = (0, co, 2 * co)
            self._conv_seg(g1, R.w1, co, out=hbuf, H=H, W=W, imgs=imgs, taps=9, seg=seg, noff=noff, bias=R.b1,
                           rowvec=rowvec, c1=R.cin, gn_ws=self._stats_for(hbuf, imgs, H * W), gn_groups=G)
            g2 = self.buf(f"{tag}.gn2", M, co)
            self._gn(hbuf, g2, R.n2g, R.n2b, imgs, H * W, True)
            okw = self._gn_kw(out, imgs, H * W)
            if R.wsc is not None:
                self._conv_seg(g2, R.w2, co, out=out, H=H, W=W, imgs=imgs, taps=9, seg=seg, noff=noff, noff2=noff,
                               bias=R.b2, a2=x, b2=R.wsc, c1=co, **okw)
            else:
                self._conv_seg(g2, R.w2, co, out=out, H=H, W=W, imgs=imgs, taps=9, seg=seg, noff=noff, bias=R.b2,
                               residual=x, c1=co, **okw)
            return out
        # unfused: t = down-conv(x) as source 2, `up` as B2 (the UNet rows take no second source)
        zero, lo2 = (0, 0, 0), (-1, 0, co)
        t1 = self._lora_t(g1, R.lora["conv1"], seg, H, W, imgs, 9, tag + ".c1", R.cin)
        self._conv_seg(g1, R.w1, co, out=hbuf, H=H, W=W, imgs=imgs, taps=9, seg=seg, noff=zero, noff2=lo2, bias=R.b1,
                       rowvec=rowvec, c1=R.cin, a2=t1, b2=R.lora["conv1"][1],
                       gn_ws=self._stats_for(hbuf, imgs, H * W), gn_groups=G)
        g2 = self.buf(f"{tag}.gn2", M, co)
        self._gn(hbuf, g2, R.n2g, R.n2b, imgs, H * W, True)
        res = x
        if R.wsc is not None:  # the second source is taken by the LoRA: the 1x1 shortcut runs as its own GEMM
            res = self.buf(f"{tag}.sc", M, co)
            tsc = self._lora_t(x, R.lora["conv_shortcut"], seg, H, W, imgs, 1, tag + ".sc", R.cin)
            self._conv_seg(x, R.wsc, co, out=res, H=H, W=W, imgs=imgs, taps=1, seg=seg, noff=zero, noff2=lo2,
                           bias=R.bsc, a2=tsc, b2=R.lora["conv_shortcut"][1])
        t2 = self._lora_t(g2, R.lora["conv2"], seg, H, W, imgs, 9, tag + ".c2", co)
        okw = self._gn_kw(out, imgs, H * W)
        self._conv_seg(g2, R.w2, co, out=out, H=H, W=W, imgs=imgs, taps=9, seg=seg, noff=zero, noff2=lo2, bias=R.b2,
                       residual=res, c1=co, a2=t2, b2=R.lora["conv2"][1], **okw)
        return out

    def _transformer(self, T: Tfm, x, imgs, H, W, ctx, out, tag, seg):
        c, hw, nt = T.c, H * W, self.n_text
        M = imgs * hw
        heads = self.cfg.num_heads
        g = self.buf(f"{tag}.tgn", M, c)
        self._gn(x, g, T.ng, T.nb, imgs, hw, False, eps=1e-6)
        hcur = self.buf(f"{tag}.th", M, c)
        fold = self.fold_ln
        st = [self._ln_slot(M) for _ in range(3)] if fold else [None] * 3
        self._lin(T.proj_in, g, hcur, hw, seg, tag + ".pin", rowstat_out=st[0])
        ln = hcur if fold else self.buf(f"{tag}.ln", M, c)  # folded: the projections read the raw hidden state
        att = self.buf(f"{tag}.att", M, c)
        # self-attention
        if not fold:
            ops.layernorm(hcur, ln, *T.ln1)
        qkv = self.buf(f"{tag}.qkv", M, 3 * c)
        self._lin(T.qkv, ln, qkv, hw, seg, tag + ".qkv", ln_stat=st[0])
        ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], att, imgs, heads, hw, hw)
        self._lin(T.o1, att, hcur, hw, seg, tag + ".o", residual=hcur, rowstat_out=st[1])
        # cross-attention
        if not fold:
            ops.layernorm(hcur, ln, *T.ln2)
        q = self.buf(f"{tag}.q2", M, c)
        self._lin(T.q2, ln, q, hw, seg, tag + ".o", ln_stat=st[1])
        # text K/V projections depend only on the prompt and the weights: computed once per set_prompt()
        kv = self.buf(f"kv.{tag.split('.')[0]}.{T.uid}", imgs * nt, 2 * c)
        if self._kv_recompute:
            self._lin(T.kv2, ctx, kv, nt, seg, tag + ".kv")
        ops.attention(q, kv[:, :c], kv[:, c:], att, imgs, heads, hw, nt)
        self._lin(T.o2, att, hcur, hw, seg, tag + ".o", residual=hcur, rowstat_out=st[2])
        # feed-forward (GEGLU fused in the first GEMM's epilogue)
        if not fold:
            ops.layernorm(hcur, ln, *T.ln3)
        u = self.buf(f"{tag}.ff", M, 4 * c)
        self._lin(T.ff1, ln, u, hw, seg, tag + ".o", ln_stat=st[2], act=ACT_GEGLU)
        self._lin(T.ff2, u, hcur, hw, seg, tag + ".ff2", residual=hcur)
        self._lin(T.proj_out, hcur, out, hw, seg, tag + ".pout", residual=x, **self._gn_kw(out, imgs, hw))
        return out

    def _time_path(self, E: EncoderW, groups: Sequence[Tuple[int, int]], tag: str, ncols: Sequence[int], out=None):
        """groups: [(weight-group index, n images)] in pass order.  Returns temb [imgs_total, cols] fp32."""
        cfg, B = self.cfg, self.B
        c0, td = cfg.block_out_channels[0], cfg.time_embed_dim
        total = sum(n for _, n in groups)
        cols = max(ncols)
        temb = out if out is not None else self.buf(f"{tag}.temb", total, cols, torch.float32)
        sin = self.buf("t_sin", B, c0, torch.float32)
        r0 = 0
        for (gi, n), nc in zip(groups, ncols):
            e1 = self.buf(f"{tag}.e1.{gi}", B, td, torch.float32)
            e2 = self.buf(f"{tag}.e2.{gi}", B, td, torch.float32)
            ops.small_linear(sin, E.te1[gi][0], E.te1[gi][1], e1, silu_out=True)
            # every consumer of emb applies SiLU first (ResnetBlock2D.time_emb_proj(nonlinearity(temb))): store silu(emb)
            ops.small_linear(e1, E.te2[gi][0], E.te2[gi][1], e2, silu_out=True)
            # images of one group repeat the B timestep rows (n is a multiple of B)
            for rep in range(n // B):
                ops.small_linear(e2, E.temb_w[gi][:nc], E.temb_b[gi][:nc], temb[r0:r0 + B, :nc])
                r0 += B
        return temb

    def _encoder(self, E: EncoderW, x, imgs, temb, ctx, seg, tag, on_level=None):
        """x: conv_in output (+cond) [imgs*hw, c0].  Returns 12 skip tensors + mid.  `on_level(li, tensor)` is called
        right after residual level li (0..11 skips, 12 = mid) has been enqueued."""
        cfg = self.cfg
        skips = [x]
        notify = on_level if on_level is not None else (lambda li, t: None)
        notify(0, x)
        for i, (H, W) in enumerate(self.levels):
            c = cfg.block_out_channels[i]
            for j in range(cfg.layers_per_block):
                R = E.down_res[i][j]
                T = E.down_tfm[i][j]
                M = imgs * H * W
                if T is None:
                    out = self.buf(f"{tag}.skip{len(skips)}", M, c)
                    self._resnet(R, x, imgs, H, W, temb, out, f"{tag}.L{i}", seg)
                else:
                    r_out = self.buf(f"{tag}.L{i}.rout", M, c)
                    self._resnet(R, x, imgs, H, W, temb, r_out, f"{tag}.L{i}", seg)
                    out = self.buf(f"{tag}.skip{len(skips)}", M, c)
                    self._transformer(T, r_out, imgs, H, W, ctx, out, f"{tag}.L{i}", seg)
                x = out
                skips.append(x)
                notify(len(skips) - 1, x)
            if E.down_conv[i] is not None:
                Hn, Wn = self.levels[i + 1]
                col = self.buf(f"{tag}.L{i}.col", imgs * Hn * Wn, 9 * c)
                ops.im2col3x3(x, col, imgs, H, W, c, 2)
                out = self.buf(f"{tag}.skip{len(skips)}", imgs * Hn * Wn, c)
                self._convw(E.down_conv[i], col, out, Hn, Wn, imgs, seg, f"{tag}.L{i}.down",
                            gn_ws=self._stats_for(out, imgs, Hn * Wn), gn_groups=cfg.norm_num_groups)
                x = out
                skips.append(x)
                notify(len(skips) - 1, x)
        H, W = self.levels[-1]
        c = cfg.block_out_channels[-1]
        M = imgs * H * W
        m0 = self.buf(f"{tag}.mid0", M, c)
        self._resnet(E.mid_res[0], x, imgs, H, W, temb, m0, f"{tag}.mid", seg)
        m1 = self.buf(f"{tag}.mid1", M, c)
        self._transformer(E.mid_tfm, m0, imgs, H, W, ctx, m1, f"{tag}.mid", seg)
        mid = self.buf(f"{tag}.mid2", M, c)
        self._resnet(E.mid_res[1], m1, imgs, H, W, temb, mid, f"{tag}.mid", seg)
        notify(len(skips), mid)
        return skips, mid

    # ------------------------------------------------------------------------------------ the step
    def _run_step(self, active: Sequence[bool] = (True,) * 6, mode: str = "step", guess_mode: bool = False,
                  zero_uncond: bool = False):
        """mode 'step': full fused step -> eps_out.  mode 'residuals': stop after the merge and leave the 13
        merged residuals (what EdgeStyleMultiControlNetModel.forward returns) in self.res_out.

        The conditioning scales are read by the merge kernels from `self.cond_scale_dev` (a device vector), so one
        captured graph serves every scale.  `active[k]` = net k contributes (its scale is non-zero): the image blocks
        of inactive nets are dropped from the two batched encoder passes (control_guidance_start / end gating,
        /root/reference/model/edgestyle_pipeline.py:418-427, without spending the gated branch's FLOPs).

        guess_mode: every ControlNet scales its 13 outputs by logspace(-1, 0, 13) * conditioning_scale
        (controllora.py:257-265) instead of a uniform scale.  zero_uncond: the pipeline's guess mode under CFG
        (edgestyle_pipeline.py:453-459, 487-497) -- the merged residuals of the unconditional rows (first half of the
        batch) are dropped, i.e. those rows keep the plain UNet skips."""
        if not self.full:
            raise RuntimeError("this engine was built for a stand-alone ControlNet: the fused step needs the UNet, both "
                               "ControlLoRA nets, the openpose net and the merge blocks")
        cfg, B = self.cfg, self.B
        h, w = self.h, self.w
        hw = h * w
        boc = cfg.block_out_channels
        c0 = boc[0]
        nt = self.n_text
        active = tuple(bool(a) for a in active)
        geo = self._geometry(active)
        self._ensure_text_kv(active)
        self._begin_step_scratch()
        ops.PREFETCH.gate = self.weight_prefetch != "decoder"
        # -- sample: NCHW fp32 -> NHWC, im2col (K = 36 padded to 64)
        s16 = self.buf("sample16", B * hw, 8)
        ops.nchw_to_nhwc(self.sample_in, s16)
        col = self.buf("sample_col", B * hw, 64)
        ops.im2col3x3(s16, col, B, h, w, cfg.in_channels, 1)
        ops.timestep_embedding(self.t_in, c0, self.buf("t_sin", B, c0, torch.float32))
        Eb, Ep = self.enc_base, self.enc_pose
        enc_cols = Eb.temb_cols
        # fork here: the pose chain only needs the sinusoidal embedding and the im2col of the sample.  The base
        # chain's time path (a dozen weight-streaming GEMVs) runs on the otherwise idle merge stream while the main
        # stream does conv_in and the first GroupNorm; the first resnet waits for it just before its conv1.
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        self._merge_stream.wait_event(fork)
        with torch.cuda.stream(self._merge_stream):
            groups = [(0, B)] + ([(1, B)] if geo.seg[1] else []) + ([(2, geo.seg[2])] if geo.seg[2] else [])
            temb_base = self._time_path(Eb, groups, geo.btag,
                                        [enc_cols + self.dec_temb_cols] + [enc_cols] * (len(groups) - 1))
            self._temb_ready = torch.cuda.Event()
            self._temb_ready.record(self._merge_stream)
            self._temb_waited = set()
        cond = lambda k: self.conds[k * B * hw:(k + 1) * B * hw]
        # -- encoder passes.  Image order of the base weight set: unet (B) | agn (B) | clo cond2 (B) | clo cond4 (B)
        #    (inactive nets dropped); of the pose set: cond1 | cond3 | cond5.  The two passes run on two streams: most
        #    layers at CFG batch 2 are latency / occupancy bound, so the concurrent pass fills the SMs that one chain's
        #    partial waves leave idle.
        nb_b, nb_p = len(geo.base_nets), len(geo.pose_nets)
        xb = self.buf(f"{geo.btag}.x0", nb_b * B * hw, c0)
        done_events = []
        # conv_in (+ cached conditioning embedding as the residual, controllora.py:203): one GEMM per image block
        ci = Eb.conv_in
        for pos, k in enumerate(geo.base_nets):
            grp = 0 if k is None else (1 if k == 0 else 2)
            self._convw_block(ci, col, xb[pos * B * hw:(pos + 1) * B * hw], grp, f"{geo.btag}.ci{pos}",
                              residual=None if k is None else cond(k))
        n_lora = nb_b - 1
        zres = {}  # ("b" | "p", level) -> zero-conv output of the ControlLoRA / openpose image blocks
        zc_used = set()

        def zero_conv(kind, li, t):
            """controllora.py:240-254 for residual level li of one pass: the ControlLoRA image blocks (agn | clo | clo)
            are one GEMM with the weight set selected per row segment, the openpose blocks one GEMM."""
            c, H, W = self.res_shapes[li]
            n = B * H * W
            if kind == "b":
                rb = self.buf(f"zres_b{li}.{n_lora}", n_lora * n, c)
                n_agn = n if geo.seg[1] else 0
                zw, zb = self.zero_base[li]
                ops.gemm(t[n:], zw, c, out=rb, bias=zb, segs=([0, n_agn, n_lora * n], [0, c], None))
                zres[("b", li)] = rb
            else:
                rp = self.buf(f"zres_p{li}.{nb_p}", nb_p * n, c)
                ops.gemm(t, self.zero_pose[li][0], c, out=rp, bias=self.zero_pose[li][1])
                zres[("p", li)] = rp

        def early(kind, zst):
            """Callback of _encoder: as soon as a residual level is enqueued, its zero-conv goes to the side stream."""
            if not self.zc_early:
                return None

            def on_level(li, t):
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                zst.wait_event(ev)
                zc_used.add(zst)
                with torch.cuda.stream(zst):
                    zero_conv(kind, li, t)
            return on_level

        outs_b = self._encoder(Eb, xb, nb_b * B, temb_base, self.ctx_base[: nb_b * B * nt], geo.seg, geo.btag,
                               early("b", self._zc_streams[0]) if n_lora else None)
        outs_b = outs_b[0] + [outs_b[1]]
        outs_p = None
        if nb_p:
            st = self._chain_streams[0]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                xp = self.buf(f"{geo.ptag}.x0", nb_p * B * hw, c0)
                temb_pose = self._time_path(Ep, [(0, nb_p * B)], geo.ptag, [Ep.temb_cols])
                for pos, k in enumerate(geo.pose_nets):
                    ops.gemm(col, Ep.conv_in.w, c0, out=xp[pos * B * hw:(pos + 1) * B * hw], bias=Ep.conv_in.bias,
                             residual=cond(k))
                sk, md = self._encoder(Ep, xp, nb_p * B, temb_pose, self.ctx_pose[: nb_p * B * nt], None, geo.ptag,
                                       early("p", self._zc_streams[1]))
                outs_p = sk + [md]
                ev = torch.cuda.Event()
                ev.record(st)
                done_events.append(ev)
        # -- decoder concat buffers (x | skip) and their geometry
        rev = list(reversed(boc))
        rev_attn = list(reversed(cfg.down_has_attn))
        n_up = cfg.layers_per_block + 1
        skip_ch = [s[0] for s in self.res_shapes[:-1]]
        cats = {}
        x_ch, sidx = boc[-1], len(skip_ch) - 1
        for i in range(len(boc)):
            H, W = self.levels[len(boc) - 1 - i]
            for j in range(n_up):
                cats[(i, j)] = (self.buf(f"cat{i}.{j}", B * H * W, x_ch + skip_ch[sidx]), x_ch, sidx)
                x_ch = rev[i]
                sidx -= 1
        cat_of_skip = {v[2]: (v[0], v[1]) for v in cats.values()}
        # GroupNorm statistics of the concat inputs come from their two producers (no statistics pass in the decoder)
        cat_gn = {}
        if mode == "step" and not zero_uncond and os.environ.get("ES_CAT_GN", "1") != "0":
            for (i, j), (cbuf, xc, sidx_ij) in cats.items():
                H, W = self.levels[len(boc) - 1 - i]
                if self._register_cat(cbuf, B, H * W):
                    cat_gn[sidx_ij] = (self._cat_slot[cbuf.data_ptr()], xc)
        # -- zero convs (controllora.py:240-254) + EdgeStyle merge (edgestyle_multicontrolnet.py:160-169) on the side
        #    stream, in the order the decoder consumes them (mid, then skips 11..0).  All levels of a group go through
        #    ONE launch per merge phase; the decoder waits on one event per group.  Default: the deep levels first (the
        #    decoder starts on them), the three heavy 64x64-level merges under the decoder's latency-bound deep levels.
        level_gain = torch.logspace(-1, 0, len(self.res_shapes)).tolist() if guess_mode else [1.0] * len(self.res_shapes)
        nlev = len(self.res_shapes)
        merged = {}
        side = self._merge_stream
        ev_main = torch.cuda.Event()
        ev_main.record(main)
        level_groups = merge_level_groups(nlev, n_up, mode == "step", self.merge_by_level, self.merge_split)
        # zero-convs: each encoder pass does its own, deepest group first, on a stream that is idle once the encoders are
        # through (the ControlLoRA blocks on a side stream behind the main pass, the openpose blocks on the pose
        # stream behind its pass); the merge stream waits group by group
        zc_ready = [[] for _ in level_groups]
        if self.zc_chain != "0" and not self.zc_early:
            if n_lora:
                zst = self._zc_streams[0]
                zst.wait_event(ev_main)
                with torch.cuda.stream(zst):
                    for g, lg in enumerate(level_groups[:1] if self.zc_chain == "first" else level_groups):
                        for li in lg:
                            zero_conv("b", li, outs_b[li])
                        ev = torch.cuda.Event()
                        ev.record(zst)
                        zc_ready[g].append(ev)
            if nb_p:
                st = self._chain_streams[0]
                with torch.cuda.stream(st):
                    for g, lg in enumerate(level_groups[:1] if self.zc_chain == "first" else level_groups):
                        for li in lg:
                            zero_conv("p", li, outs_p[li])
                        ev = torch.cuda.Event()
                        ev.record(st)
                        zc_ready[g].append(ev)
        side.wait_event(ev_main)
        for ev in done_events:
            side.wait_event(ev)
        if self.zc_early:  # the zero-conv streams join the merge stream
            for zst in zc_used:
                ev = torch.cuda.Event()
                ev.record(zst)
                side.wait_event(ev)
        z_dtype = torch.float32 if (self.dtype == torch.bfloat16 or not self.merge_z16) else self.dtype
        with torch.cuda.stream(side):
            for g, lg in enumerate(level_groups):
                for ev in zc_ready[g]:
                    side.wait_event(ev)
                table = []
                for li in lg:
                    c, H, W = self.res_shapes[li]
                    n = B * H * W
                    res = [None] * 6
                    if n_lora:
                        if ("b", li) not in zres:
                            zero_conv("b", li, outs_b[li])
                        for pos, k in enumerate(geo.base_nets[1:]):
                            res[k] = zres[("b", li)][pos * n:(pos + 1) * n]
                    if nb_p:
                        if ("p", li) not in zres:
                            zero_conv("p", li, outs_p[li])
                        for pos, k in enumerate(geo.pose_nets):
                            res[k] = zres[("p", li)][pos * n:(pos + 1) * n]
                    unet_rows = outs_b[li][:n]
                    res = [r if r is not None else unet_rows for r in res]  # never read: their scale is 0
                    z = self.buf(f"merge_z{li}", n, c, z_dtype)
                    if mode == "residuals":
                        dst, skip = self.buf(f"res_out{li}", n, c), None
                    elif li < nlev - 1:
                        cbuf, xc = cat_of_skip[li]
                        dst, skip = cbuf[:, xc:], unet_rows
                    else:  # mid: becomes the x half of the first decoder concat
                        dst, skip = cats[(0, 0)][0][:, :c], unet_rows
                    gn = None
                    G = cfg.norm_num_groups
                    if mode == "step" and li < nlev - 1 and li in cat_gn:
                        (ws_c, cpg_c), xc_c = cat_gn[li]          # skip half of the concat that consumes level li
                        gn = (ws_c, G, cpg_c, xc_c)
                    elif mode == "step" and li == nlev - 1 and cats[(0, 0)][0].data_ptr() in self._cat_slot:
                        ws_c, cpg_c = self._cat_slot[cats[(0, 0)][0].data_ptr()]  # merged mid = x half of the first concat
                        gn = (ws_c, G, cpg_c, 0)
                    table.append(dict(res=res, prm=self.merge[li], stats=self._merge_slot(), z=z, hw=H * W, C=c, dst=dst,
                                      skip=skip, gn=gn, gain=level_gain[li]))
                ops.merge_levels(table, self.cond_scale_dev, B)
                if zero_uncond:
                    for lv in table:
                        nu = (B // 2) * lv["hw"]  # rows of the unconditional images (negative prompt rows come first)
                        if lv["skip"] is not None:
                            lv["dst"][:nu].copy_(lv["skip"][:nu])
                        else:
                            lv["dst"][:nu].zero_()
                ev = torch.cuda.Event()
                ev.record(side)
                for li in lg:
                    merged[li] = ev
        order = [li for lg in level_groups for li in lg]
        self._temb_ready = None
        if mode == "residuals":
            # the side stream is in order: the last event covers all levels
            main.wait_event(merged[order[-1]])
            return
        main.wait_event(merged[len(self.res_shapes) - 1])
        # -- UNet decoder (ES_WEIGHT_PREFETCH=decoder: only here, where one stream runs alone and HBM is idle between the
        #    layers, does a GEMM pull the next layer's weights into L2)
        ops.PREFETCH.gate = True
        for i in range(len(boc)):
            H, W = self.levels[len(boc) - 1 - i]
            M = B * H * W
            cout = rev[i]
            for j in range(n_up):
                cbuf, _, sidx_ij = cats[(i, j)]
                main.wait_event(merged[sidx_ij])
                last = (j == n_up - 1)
                if not last:
                    dest = cats[(i, j + 1)][0][:, :cout]
                elif i < len(boc) - 1:
                    dest = self.buf(f"dec{i}.out", M, cout)
                else:
                    dest = self.buf("dec.final", M, cout)
                R, T = self.up_res[i][j], self.up_tfm[i][j]
                if T is None:
                    self._resnet(R, cbuf, B, H, W, temb_base, dest, f"dec.L{i}")
                else:
                    r_out = self.buf(f"dec.L{i}.rout", M, cout)
                    self._resnet(R, cbuf, B, H, W, temb_base, r_out, f"dec.L{i}")
                    self._transformer(T, r_out, B, H, W, self.ctx_dec, dest, f"dec.L{i}", None)
            if self.up_conv[i] is not None:
                Hn, Wn = self.levels[len(boc) - 2 - i]
                assert (Hn, Wn) == (2 * H, 2 * W), "odd latent sizes are not supported by the x2 upsample path"
                up = self.buf(f"dec{i}.up", B * Hn * Wn, cout)
                ops.upsample2x(self.buf(f"dec{i}.out", M, cout), up, B, H, W)
                up_out = cats[(i + 1, 0)][0][:, :cout]
                ops.gemm(up, self.up_conv[i][0], cout, out=up_out, taps=9, whn=(Wn, Hn, B),
                         bias=self.up_conv[i][1], c1=cout, **self._gn_kw(up_out, B, Hn * Wn))
        # -- conv_norm_out + SiLU + conv_out
        fin = self.buf("dec.final", B * hw, c0)
        g = self.buf("dec.gn_out", B * hw, c0)
        self._gn(fin, g, self.norm_out[0], self.norm_out[1], B, hw, True)
        o = self.buf("dec.conv_out", B * hw, 16, torch.float32)
        ops.gemm(g, self.conv_out[0], cfg.out_channels, out=o, taps=9, whn=(w, h, B), bias=self.conv_out[1], c1=c0,
                 block_n=32)
        self._nhwc32_to_nchw(o, self.eps_out)
        ops.PREFETCH.gate = self.weight_prefetch != "decoder"

    def _nhwc32_to_nchw(self, o, dst):
        # [B*hw, 16] fp32 (first out_channels valid) -> [B, c, h, w]; tiny (64 KB): a strided copy kernel of torch
        B, c = dst.shape[0], dst.shape[1]
        dst.copy_(o.view(B, self.h * self.w, 16)[:, :, :c].permute(0, 2, 1).reshape(dst.shape))

    # ------------------------------------------------------------------------------------ public
    @torch.no_grad()
    def step(self, sample: torch.Tensor, timestep, cond_scale: Sequence[float] = (1.0,) * 6, guess_mode: bool = False,
             zero_uncond: bool = False) -> torch.Tensor:
        """noise_pred = UNet(sample, t, ehs, residuals(6 ControlNets + merge)) -- the fused single-step form the
        reference defines at /root/reference/export_onnx.py:43-74.  Returns a view of the static output buffer."""
        self._load_sample_t(sample, timestep)
        key = self._set_cond_scale(cond_scale)
        if guess_mode or zero_uncond:  # the rare path: eager (no graph per flag combination)
            n0 = ops.LAUNCHES
            self._run_step(key, guess_mode=guess_mode, zero_uncond=zero_uncond)
            self.launches_per_step = ops.LAUNCHES - n0
            return self.eps_out
        if not self.use_graph:
            n0 = ops.LAUNCHES
            self._run_step(key)
            self.launches_per_step = ops.LAUNCHES - n0
            self._finish_tuning()
            return self.eps_out
        gph = self._graphs.get(key)
        if gph is None:
            # one graph per SET of contributing nets (not per scale value: the scales are a device vector)
            if self._tune:
                ops.TUNER.enabled = True
                self._tuning_done = False
            self._run_step(key)  # eager warm-up: allocates every buffer, sets kernel attributes, tunes GEMM tiles
            torch.cuda.synchronize()
            self._finish_tuning()
            # next-layer weight prefetch into L2 (ops.WeightPrefetchPlan): measured 11.55 vs 11.40 ms/step -- the extra
            # HBM stream competes with the running layer's own operand traffic -- so it is off unless asked for
            prefetch = self.weight_prefetch != "0"
            if prefetch:  # second eager pass (tuning is frozen now): record the per-stream order of the weight tensors
                ops.PREFETCH.begin("record")
                self._run_step(key)
                torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            if prefetch:
                ops.PREFETCH.begin("use")
            try:
                with torch.cuda.graph(gph):
                    self._run_step(key)
            finally:
                ops.PREFETCH.end()
            self._graph_launches[key] = ops.LAUNCHES - n0
            self._graphs[key] = gph
        self.launches_per_step = self._graph_launches[key]
        gph.replay()
        return self.eps_out

    def _set_cond_scale(self, cond_scale: Sequence[float]) -> Tuple[bool, ...]:
        """Upload the six conditioning scales (only when they changed) and return the set of contributing nets: the
        key of the step's schedule / CUDA graph.  ES_SKIP_GATED=0 keeps gated nets in the batched passes (their
        residuals are then multiplied by 0 inside the merge, as in the reference)."""
        sc = tuple(float(x) for x in cond_scale)
        assert len(sc) == 6
        if sc != self._scale_host:
            self.cond_scale_dev.copy_(torch.tensor(sc, dtype=torch.float32), non_blocking=False)
            self._scale_host = sc
        if os.environ.get("ES_SKIP_GATED", "1") == "0":
            return (True,) * 6
        return tuple(x != 0.0 for x in sc)

    def _finish_tuning(self):
        """Called after the first full eager step: freeze the tuner (lookups stay active) and dump the table."""
        if ops.TUNER.enabled and not self._tuning_done:
            ops.TUNER.enabled = False
            self._tuning_done = True
            dump = os.environ.get("ES_TUNE_DUMP")
            if dump:
                ops.TUNER.save(dump)

    def _load_sample_t(self, sample, timestep):
        self.sample_in.copy_(sample.to(device=self.dev, dtype=torch.float32))
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)])
        t = timestep.to(device=self.dev, dtype=torch.float32).reshape(-1)
        self.t_in.copy_(t.expand(self.B) if t.numel() == 1 else t)

    @torch.no_grad()
    def residuals(self, sample, timestep, cond_scale: Sequence[float],
                  guess_mode: bool = False) -> Tuple[List[torch.Tensor], torch.Tensor]:
        """EdgeStyleMultiControlNetModel.forward (edgestyle_multicontrolnet.py:116-171): 12 merged down residuals +
        mid as fresh NCHW fp32 tensors."""
        self._load_sample_t(sample, timestep)
        self._run_step(self._set_cond_scale(cond_scale), mode="residuals", guess_mode=guess_mode)
        outs = []
        for li, (c, H, W) in enumerate(self.res_shapes):
            dst = torch.empty(self.B, c, H, W, device=self.dev, dtype=torch.float32)
            ops.nhwc_to_nchw(self.buf(f"res_out{li}", self.B * H * W, c), dst)
            outs.append(dst)
        return outs[:-1], outs[-1]

    @torch.no_grad()
    def single_controlnet(self, group: Optional[int], sample, timestep, prompt_embeds, cond, conditioning_scale: float,
                          guess_mode: bool = False) -> Tuple[List[torch.Tensor], torch.Tensor]:
        """CachedControlNetModel.forward (controllora.py:59-287) for ONE net: group 0/1 = ControlLoRA sets (UNet base
        weights + that LoRA), None = the plain (openpose) ControlNet.  Returns NCHW fp32 residuals x scale."""
        cfg, B, h, w = self.cfg, self.B, self.h, self.w
        hw, c0, nt = h * w, cfg.block_out_channels[0], self.n_text
        self._load_sample_t(sample, timestep)
        self._begin_step_scratch()
        s16 = self.buf("sample16", B * hw, 8)
        ops.nchw_to_nhwc(self.sample_in, s16)
        col = self.buf("sample_col", B * hw, 64)
        ops.im2col3x3(s16, col, B, h, w, cfg.in_channels, 1)
        ops.timestep_embedding(self.t_in, c0, self.buf("t_sin", B, c0, torch.float32))
        E = self.enc_pose if group is None else self.enc_base
        gi = 0 if group is None else group + 1
        temb = self._time_path(E, [(gi, B)], "single", [E.temb_cols])
        cond16 = self.buf("single.cond", B * hw, c0)
        ops.nchw_to_nhwc(cond.to(device=self.dev, dtype=torch.float32).contiguous(), cond16)
        x0 = self.buf("single.x0", B * hw, c0)
        self._convw_block(E.conv_in, col, x0, 0 if group is None else group + 1, "single.ci", residual=cond16)
        ctx = self.buf("single.ctx", B * nt, cfg.cross_attention_dim)
        ctx.copy_(prompt_embeds.to(device=self.dev, dtype=self.dtype).reshape(B * nt, -1))
        seg = None if group is None else ((0, B, 0) if group == 0 else (0, 0, B))
        self._kv_recompute = True  # the prompt is an argument of this call: no cached text K/V
        try:
            skips, mid = self._encoder(E, x0, B, temb, ctx, seg, "single")
        finally:
            self._kv_recompute = False
        n_res = len(self.res_shapes)
        if guess_mode:  # controllora.py:257-265
            scales = (torch.logspace(-1, 0, n_res) * conditioning_scale).tolist()
        else:
            scales = [conditioning_scale] * n_res
        outs = []
        for li, ((c, H, W), src) in enumerate(zip(self.res_shapes, skips + [mid])):
            n = B * H * W
            r = self.buf(f"single.z{li}", n, c)
            if group is None:
                ops.gemm(src, self.zero_pose[li][0], c, out=r, bias=self.zero_pose[li][1], alpha=scales[li])
            else:
                zw, zb = self.zero_base[li]
                ops.gemm(src, zw, c, out=r, bias=zb, alpha=scales[li], segs=([0, n], [group * c], None))
            dst = torch.empty(B, c, H, W, device=self.dev, dtype=torch.float32)
            ops.nhwc_to_nchw(r, dst)
            outs.append(dst)
        return outs[:-1], outs[-1]

    @torch.no_grad()
    def embed_openpose(self, image: torch.Tensor) -> torch.Tensor:
        """ControlNetConditioningEmbedding of the openpose ControlNet: image [n, 3, 8h', 8w'] (fp32, NCHW, values as the
        reference's image processor produces them) -> conditioning embedding [n, c0, h', w'] fp32 NCHW.  Runs once per
        pipeline call (prepare_image, edgestyle_pipeline.py:629-664), not per step."""
        if self.pose_embed is None:
            raise ValueError("the openpose checkpoint carries no controlnet_cond_embedding weights")
        n, cin, H, W = image.shape
        x = self.buf(f"pemb.in.{n}x{H}x{W}", n * H * W, self.pose_embed[0][2])
        ops.nchw_to_nhwc(image.to(device=self.dev, dtype=torch.float32).contiguous(), x)
        for li, (wgt, bias, cin_pad, cout, stride, silu) in enumerate(self.pose_embed):
            assert x.shape[1] == cin_pad, (x.shape, cin_pad)
            act = ACT_SILU if silu else 0
            if stride == 1:
                out = self.buf(f"pemb.{li}.{n}x{H}x{W}", n * H * W, cout)
                ops.gemm(x, wgt, cout, out=out, taps=9, whn=(W, H, n), bias=bias, c1=cin_pad, act=act)
            else:
                Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
                col = self.buf(f"pemb.col{li}.{n}x{H}x{W}", n * Ho * Wo, 9 * cin_pad)
                ops.im2col3x3(x, col, n, H, W, cin_pad, 2)
                out = self.buf(f"pemb.{li}.{n}x{H}x{W}", n * Ho * Wo, cout)
                ops.gemm(col, wgt, cout, out=out, bias=bias, act=act)
                H, W = Ho, Wo
            x = out
        dst = torch.empty(n, x.shape[1], H, W, device=self.dev, dtype=torch.float32)
        ops.nhwc_to_nchw(x, dst)
        return dst

    @torch.no_grad()
    def embed_vae_latent(self, z: torch.Tensor, group: Optional[int] = None) -> torch.Tensor:
        """Tail of VAEControlNetConditioningEmbedding.forward (controllora.py:40-41): `conv_vae_out(z)` where z is the
        scaled VAE latent [n, 4, h, w] fp32 NCHW.  `conv_vae_out` IS the net's `conv_in` module (:36), whose parameters
        `tie_weights` points at the UNet's conv_in (:624), so this runs the UNet conv_in weights (SURVEY.md appendix,
        quirk 1) -> [n, c0, h, w] fp32 NCHW.  Once per pipeline call, not per step."""
        cfg = self.cfg
        n, c, h, w = z.shape
        assert c == cfg.in_channels, (c, cfg.in_channels)
        c0 = cfg.block_out_channels[0]
        s16 = self.buf(f"vemb.in.{n}x{h}x{w}", n * h * w, 8)
        ops.nchw_to_nhwc(z.to(device=self.dev, dtype=torch.float32).contiguous(), s16)
        col = self.buf(f"vemb.col.{n}x{h}x{w}", n * h * w, 64)
        ops.im2col3x3(s16, col, n, h, w, cfg.in_channels, 1)
        out = self.buf(f"vemb.out.{n}x{h}x{w}", n * h * w, c0)
        # `group`: the ControlLoRA net the embedder belongs to -- with a conv LoRA (lora_conv2d_rank > 0) its conv_in
        # carries that net's low-rank update, and conv_vae_out is that very module
        self._convw_block(self.enc_base.conv_in, col, out, 0 if group is None else group + 1, "vemb.ci")
        dst = torch.empty(n, c0, h, w, device=self.dev, dtype=torch.float32)
        ops.nhwc_to_nchw(out, dst)
        return dst

    @torch.no_grad()
    def cfg_ddim_update(self, latents: torch.Tensor, a_t: float, a_prev: float, guidance=None):
        """CFG combine (edgestyle_pipeline.py:513-517) + DDIM update (:520-522) on eps_out, in place on `latents`."""
        import math

        self.coef.copy_(torch.tensor([math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_prev), math.sqrt(1 - a_prev)]))
        if guidance is not None:
            g = guidance if torch.is_tensor(guidance) else torch.full((self.B // 2,), float(guidance))
            self.guidance.copy_(g.to(torch.float32).reshape(-1).expand(self.B // 2))
        ops.cfg_ddim(self.eps_out, latents, self.guidance, self.coef)
        return latents
