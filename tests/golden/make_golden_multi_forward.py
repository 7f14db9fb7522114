"""Golden vector for `EdgeStyleMultiControlNetModel.forward` from the reference's OWN source text (build container only).

The class subclasses diffusers' MultiControlNetModel and cannot be imported, but its `forward`
(/root/reference/model/edgestyle_multicontrolnet.py:116-171) is pure torch around `self.nets`,
`self.multi_controlnet_down_blocks` and `self.multi_controlnet_mid_block`.  This script exec()s that method together
with `ControlNetBlock` / `interleave_tensors*` (nothing is copied into the repo) on a stub `self`: six stub nets whose
13 outputs depend on THEIR conditioning image and scale (so the routing of images / scales to nets, the level-wise
zip, the channel interleave and the per-level blocks are all pinned), at the SD1.5 residual pattern scaled down
(channels 8/16/32/32, 8x8 latent).  Inputs, block parameters and outputs go to `multi_forward_golden.pt`, replayed
against `oracle.merge.EdgeStyleMultiControlNetModel` by `tests/test_oracle_golden.py`.

    python tests/golden/make_golden_multi_forward.py
"""
import ast
import os
import sys
import types
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

REF = "/root/reference/model/edgestyle_multicontrolnet.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_forward_golden.pt")
CH = [8, 8, 8, 8, 16, 16, 16, 32, 32, 32, 32, 32]
HW = [8, 8, 8, 4, 4, 4, 2, 2, 2, 1, 1, 1]


def load_reference_symbols():
    src = open(REF).read()
    ns = {"torch": torch, "nn": nn, "List": List, "Tuple": Tuple, "Union": Union, "Optional": Optional, "Dict": Dict,
          "Any": Any, "Callable": Callable, "ControlNetOutput": object}
    for node in ast.parse(src).body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in (
                "ControlNetBlock", "interleave_tensors", "interleave_tensors_from_list_of_lists", "zero_module",
                "ones_module"):
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "EdgeStyleMultiControlNetModel":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "forward":
                    import textwrap

                    exec(compile(textwrap.dedent(ast.get_source_segment(src, sub)), REF, "exec"), ns)
    return ns


def stub_outputs(base, image, scale):
    """What stub net k returns: its fixed base tensors shifted by its conditioning image's mean, times its scale."""
    shift = image.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)
    outs = [(b + shift) * scale for b in base]
    return outs[:-1], outs[-1]


def main():
    ns = load_reference_symbols()
    Block, forward = ns["ControlNetBlock"], ns["forward"]
    g = torch.Generator().manual_seed(20240609)
    B = 2
    torch.manual_seed(5)
    blocks = [Block(c, (s, s), 6) for c, s in zip(CH, HW)] + [Block(32, (1, 1), 6)]
    with torch.no_grad():
        for blk in blocks:
            for p in blk.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
    bases = [[torch.randn(B, c, s, s, generator=g) for c, s in zip(CH + [32], HW + [1])] for _ in range(6)]
    images = [torch.randn(B, 8, 8, 8, generator=g) for _ in range(6)]
    scales = [1.0, 0.5, 2.0, 1.5, 0.0, 0.75]

    def make_net(k):
        def net(sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale, **kw):
            return stub_outputs(bases[k], controlnet_cond, conditioning_scale)
        return net

    self = types.SimpleNamespace(nets=[make_net(k) for k in range(6)], multi_controlnet_down_blocks=blocks[:-1],
                                 multi_controlnet_mid_block=blocks[-1])
    with torch.no_grad():
        down, mid = forward(self, torch.zeros(B, 4, 8, 8), torch.tensor(1), torch.zeros(B, 77, 8), images, scales)
    assert len(down) == 12
    torch.save({"ch": CH, "hw": HW, "blocks": [{k: v.clone() for k, v in b.state_dict().items()} for b in blocks],
                "bases": bases, "images": images, "scales": scales, "down": list(down), "mid": mid}, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
