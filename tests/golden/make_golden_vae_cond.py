"""Golden vector for the VAE conditioning embedder from the reference's OWN source text (build container only).

/root/reference/model/controllora.py imports diffusers at module scope and cannot be imported, but
`VAEControlNetConditioningEmbedding` (:28-42) and `_tie_weights` (:45-56) are pure torch around two names:
`zero_module` (diffusers.models.controlnet: zero every parameter, return the module -- re-stated here) and the
`autoencoder` object, for which the oracle's AutoencoderKL stands in (`encode(x).latent_dist.sample()`,
`config.scaling_factor`).  The script exec()s exactly those two source spans (nothing is copied into the repo),
builds the embedder on a seeded tiny VAE, ties `conv_vae_out` to a seeded "UNet conv_in" like
`ControlLoRAModel.tie_weights` (:623-624) does, and stores inputs + parameters + RNG seed + output in
`vae_cond_golden.pt`, which `tests/test_oracle_golden.py` replays against `oracle.controllora`.

    python tests/golden/make_golden_vae_cond.py
"""
import ast
import os
import sys
from typing import Optional

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/model/controllora.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vae_cond_golden.pt")


def zero_module(module):  # diffusers.models.controlnet.zero_module
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


def load_reference_symbols():
    src = open(REF).read()
    ns = {"torch": torch, "nn": nn, "Optional": Optional, "AutoencoderKL": object, "zero_module": zero_module}
    for node in ast.parse(src).body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in ("VAEControlNetConditioningEmbedding",
                                                                               "_tie_weights"):
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
    return ns


def main():
    from oracle.vae import AutoencoderKL, VaeConfig

    ns = load_reference_symbols()
    Emb, tie = ns["VAEControlNetConditioningEmbedding"], ns["_tie_weights"]
    torch.manual_seed(20240608)
    chans = (32, 32, 32, 32)
    vae = AutoencoderKL(VaeConfig(block_out_channels=chans, layers_per_block=1)).eval()
    with torch.no_grad():
        for k, p in vae.named_parameters():
            if "norm" in k or k.endswith(".bias"):
                p.add_(0.1 * torch.randn_like(p))
    conv_in = nn.Conv2d(4, 16, 3, padding=1)          # the ControlLoRA net's conv_in module
    emb = Emb(conv_unet=conv_in, autoencoder=vae)     # :36 zeroes it and registers it as conv_vae_out
    assert emb.conv_vae_out is conv_in and float(conv_in.weight.detach().abs().max()) == 0.0
    unet_conv_in = nn.Conv2d(4, 16, 3, padding=1)     # the UNet's conv_in
    tie(unet_conv_in, conv_in)                        # :623-624 (tie_weights)
    assert emb.conv_vae_out.weight is unet_conv_in.weight
    image = torch.rand(2, 3, 32, 32) * 2 - 1
    seed = 777
    torch.manual_seed(seed)
    with torch.no_grad():
        out = emb(image)
    torch.save({"chans": chans, "vae": {k: v.clone() for k, v in vae.state_dict().items()},
                "conv_in": {k: v.clone() for k, v in unet_conv_in.state_dict().items()}, "image": image, "seed": seed,
                "out": out}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes", tuple(out.shape))


if __name__ == "__main__":
    sys.exit(main())
