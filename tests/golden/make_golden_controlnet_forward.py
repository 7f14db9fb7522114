"""Golden vector for `CachedControlNetModel.forward` from the reference's OWN source text (build container only).

The class subclasses diffusers' ControlNetModel and cannot be imported, but its `forward`
(/root/reference/model/controllora.py:59-287) is pure torch around the sub-modules it calls.  This script exec()s
that method (nothing is copied into the repo) on a stub `self` whose sub-modules are the ORACLE's ControlNet parts
behind thin keyword adapters (diffusers calling convention -> oracle calling convention), so what gets pinned is
everything the reference itself owns on this path: timestep handling (python number / 0-d / batched tensor), conv_in +
conditioning (embedder skipped for latent-sized conditioning, :199-203), the order of the 12 skip tensors, zero-convs,
conditioning_scale and the guess_mode logspace gains (:257-270).  Weights are refilled from a seeded CPU generator in
named_parameters order (same routine in the replaying test), so only inputs' seeds and outputs are stored in
`controlnet_forward_golden.pt`.

    python tests/golden/make_golden_controlnet_forward.py
"""
import ast
import os
import sys
import textwrap
import types
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/model/controllora.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "controlnet_forward_golden.pt")
CFG = dict(block_out_channels=(16, 32, 32, 32), cross_attention_dim=16, norm_num_groups=8,
           conditioning_embedding_out_channels=(8, 8, 16, 16))


def seeded_fill(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if "norm" in name and name.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.08 * torch.randn(p.shape, generator=g))


def cases():
    g = torch.Generator().manual_seed(4242)
    B = 2
    sample = torch.randn(B, 4, 8, 8, generator=g)
    ehs = torch.randn(B, 7, 16, generator=g)
    cond_latent = torch.randn(B, 16, 8, 8, generator=g)
    cond_image = torch.rand(B, 3, 64, 64, generator=g)
    return [dict(sample=sample, timestep=951, ehs=ehs, cond=cond_latent, scale=1.0, guess_mode=False),
            dict(sample=sample, timestep=torch.tensor(17), ehs=ehs, cond=cond_image, scale=0.6, guess_mode=False),
            dict(sample=sample, timestep=torch.tensor([300, 5]), ehs=ehs, cond=cond_latent, scale=1.5, guess_mode=True),
            dict(sample=sample, timestep=12.5, ehs=ehs, cond=cond_latent, scale=0.0, guess_mode=False)]


def main():
    from oracle.sd15 import ControlNetModel, SD15Config, timestep_sinusoid

    src = open(REF).read()
    ns = {"torch": torch, "Any": Any, "Dict": Dict, "List": List, "Optional": Optional, "Tuple": Tuple, "Union": Union,
          "ControlNetOutput": lambda down_block_res_samples, mid_block_res_sample: (down_block_res_samples,
                                                                                   mid_block_res_sample)}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "CachedControlNetModel":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "forward":
                    exec(compile(textwrap.dedent(ast.get_source_segment(src, sub)), REF, "exec"), ns)
    forward = ns["forward"]
    cfg = SD15Config(**CFG)
    net = ControlNetModel(cfg).eval()
    seeded_fill(net, 99)

    class Down:
        def __init__(self, blk):
            self.blk, self.has_cross_attention = blk, blk.has_cross_attention

        def __call__(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None,
                     cross_attention_kwargs=None):
            assert attention_mask is None and cross_attention_kwargs is None
            x, outs = self.blk(hidden_states, temb, encoder_hidden_states)
            return x, tuple(outs)

    class Mid:
        has_cross_attention = True

        def __call__(self, sample, emb, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None):
            return net.mid_block(sample, emb, encoder_hidden_states)

    stub = types.SimpleNamespace(
        config=types.SimpleNamespace(controlnet_conditioning_channel_order="rgb", class_embed_type=None,
                                     addition_embed_type=None, global_pool_conditions=False),
        time_proj=lambda t: timestep_sinusoid(t, cfg.block_out_channels[0]),
        time_embedding=lambda t_emb, timestep_cond: net.time_embedding(t_emb), class_embedding=None,
        conv_in=net.conv_in, controlnet_cond_embedding=net.controlnet_cond_embedding,
        down_blocks=[Down(b) for b in net.down_blocks], mid_block=Mid(),
        controlnet_down_blocks=net.controlnet_down_blocks, controlnet_mid_block=net.controlnet_mid_block,
        dtype=torch.float32)
    outs = []
    with torch.no_grad():
        for c in cases():
            down, mid = forward(stub, c["sample"], c["timestep"], c["ehs"], c["cond"], conditioning_scale=c["scale"],
                                guess_mode=c["guess_mode"], return_dict=False)
            assert len(down) == 12
            outs.append({"down": [d.clone() for d in down], "mid": mid.clone()})
    torch.save({"cfg": CFG, "weight_seed": 99, "outs": outs}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
