"""Golden vectors from an INDEPENDENT port of diffusers modules that exists on this box: TVM's relax frontend carries
ports of HF diffusers' `get_timestep_embedding`, `Timesteps`, `TimestepEmbedding` and `Attention`
(tilelang/3rdparty/tvm/python/tvm/relax/frontend/nn/op.py:1741-1804 and modules.py:750-960).  TVM itself is
not importable here (no native runtime), so -- exactly as the other make_golden*.py scripts do for the reference -- the
SOURCE TEXT of those spans is executed over small torch-backed stand-ins for the relax operators they call
(`astype/arange/exp/expand_dims/concat/sin/cos/pad`, `reshape` with TVM's 0 = "copy this dim", `Linear`, `SiLU`,
`scaled_dot_product_attention` over [batch, seq, heads, dim], which is relax's documented layout).

    python tests/golden/make_golden_tvm_port.py   ->  tests/golden/tvm_port_golden.pt

The oracle's time embedding and Attention (oracle/sd15.py, "parity unpinned" for diffusers internals) are replayed
against these vectors by tests/test_oracle_golden.py::test_tvm_port_*.
"""
from __future__ import annotations

import math
import os
import re
import types
from typing import Optional

import torch

NN = "/opt/prime-rl/.venv/lib/python3.12/site-packages/tilelang/3rdparty/tvm/python/tvm/relax/frontend/nn"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tvm_port_golden.pt")


def _span(path, start_pat, end_pat):
    src = open(path).read()
    a = re.search(start_pat, src, re.M).start()
    m = re.search(end_pat, src[a + 1:], re.M)
    return src[a:(m.start() + a + 1) if m else len(src)]


class Tensor:  # relax nn.Tensor stand-in: `._expr` is what the raw operators consume
    def __init__(self, t):
        self._expr = t


def _raw(x):
    return x._expr if isinstance(x, Tensor) else x


def _make_env():
    # ---- raw relax operators (`_op.*`) over torch tensors
    _op = types.SimpleNamespace(
        astype=lambda x, dt: _raw(x).to(getattr(torch, dt)),
        arange=lambda start, end, dtype: torch.arange(start, end, dtype=getattr(torch, dtype)),
        exp=torch.exp, sin=torch.sin, cos=torch.cos,
        expand_dims=lambda x, axis: _raw(x).unsqueeze(axis),
        concat=lambda xs, axis: torch.cat(list(xs), dim=axis),
        nn=types.SimpleNamespace(pad=lambda x, pw: torch.nn.functional.pad(x, (pw[2], pw[3]) if len(pw) == 4 else pw)),
    )
    rx = types.SimpleNamespace(const=lambda v, dt: torch.tensor(v, dtype=getattr(torch, dt)))

    def reshape(x, shape):  # TVM reshape: 0 copies the input dim at that position, -1 infers
        t = _raw(x)
        shp = [t.shape[i] if s == 0 else s for i, s in enumerate(shape)]
        return Tensor(t.reshape(shp))

    def sdpa(q, k, v, is_causal=False):  # relax layout: [batch, seq, heads, head_dim]
        o = torch.nn.functional.scaled_dot_product_attention(_raw(q).transpose(1, 2), _raw(k).transpose(1, 2),
                                                             _raw(v).transpose(1, 2), is_causal=is_causal)
        return Tensor(o.transpose(1, 2))

    env = dict(math=math, Optional=Optional, Tensor=Tensor, _op=_op, rx=rx, get_default_dtype=lambda: "float32",
               wrap_nested=lambda e, name: Tensor(e))
    exec(_span(os.path.join(NN, "op.py"), r"^def get_timestep_embedding\(", r"^def "), env)
    op = types.SimpleNamespace(get_timestep_embedding=env["get_timestep_embedding"], reshape=reshape,
                               scaled_dot_product_attention=sdpa)

    class Module:
        def __call__(self, *a, **k):
            return self.forward(*a, **k)

    class Linear(Module):
        def __init__(self, i, o, bias=True):
            g = torch.Generator().manual_seed(i * 1000 + o + (1 if bias else 0))
            self.weight = torch.randn(o, i, generator=g) * i ** -0.5
            self.bias = torch.randn(o, generator=g) * 0.1 if bias else None

        def forward(self, x):
            return Tensor(torch.nn.functional.linear(_raw(x), self.weight, self.bias))

    class SiLU(Module):
        def forward(self, x):
            return Tensor(torch.nn.functional.silu(_raw(x)))

    class ModuleList(list):
        pass

    menv = dict(Optional=Optional, Tensor=Tensor, Module=Module, Linear=Linear, SiLU=SiLU, ModuleList=ModuleList,
                GroupNorm=None, op=op)
    exec(_span(os.path.join(NN, "modules.py"), r"^class TimestepEmbedding\(Module\):", r"^class (?!TimestepEmbedding|Timesteps|Attention)"),
         menv)
    return env, menv


def main():
    env, menv = _make_env()
    out = {"source": "tvm relax frontend nn (op.py:1741-1804, modules.py:750-960), executed from source text"}
    # 1. sinusoidal timestep embedding with the SD1.5 arguments (flip_sin_to_cos=True, freq_shift=0)
    t = torch.tensor([1.0, 51.0, 501.0, 951.0, 999.0])
    ts = menv["Timesteps"](320, flip_sin_to_cos=True, downscale_freq_shift=0)
    out["timesteps"] = {"t": t, "dim": 320, "emb": ts(Tensor(t))._expr}
    # 2. TimestepEmbedding: linear_1 -> SiLU -> linear_2
    te = menv["TimestepEmbedding"](320, 256)  # (reduced width: the formula is width-agnostic, the fixture stays small)
    x = out["timesteps"]["emb"]
    out["time_embedding"] = {"x": x, "w1": te.linear_1.weight, "b1": te.linear_1.bias, "w2": te.linear_2.weight,
                             "b2": te.linear_2.bias, "y": te(Tensor(x))._expr}
    # 3. Attention: self (dim 160, 8 heads of 20) and cross (context 96, 77 tokens)
    g = torch.Generator().manual_seed(3)
    for name, ctx_dim, n_ctx in (("self_attention", None, None), ("cross_attention", 96, 77)):
        att = menv["Attention"](160, ctx_dim, heads=8, dim_head=20)
        hs = torch.randn(2, 64, 160, generator=g)
        ehs = None if ctx_dim is None else torch.randn(2, n_ctx, ctx_dim, generator=g)
        y = att(Tensor(hs), None if ehs is None else Tensor(ehs))._expr
        out[name] = {"hidden_states": hs, "encoder_hidden_states": ehs, "to_q": att.to_q.weight, "to_k": att.to_k.weight,
                     "to_v": att.to_v.weight, "to_out_w": att.to_out[0].weight, "to_out_b": att.to_out[0].bias,
                     "qkv_bias": att.to_q.bias is not None, "scale": att.scale, "y": y}
    torch.save(out, OUT)
    print("wrote", OUT, {k: (tuple(v["y"].shape) if isinstance(v, dict) and "y" in v else None) for k, v in out.items()})


if __name__ == "__main__":
    main()
