"""Golden checkpoint directory written by the reference's OWN `save_pretrained` (build container only).

`EdgeStyleMultiControlNetModel.state_dict / load_state_dict / save_pretrained`
(/root/reference/model/edgestyle_multicontrolnet.py:173-282) are pure torch + safetensors around `self.nets`, the merge
blocks and three diffusers names (`SAFETENSORS_WEIGHTS_NAME` = "diffusion_pytorch_model.safetensors", `WEIGHTS_NAME`,
`_add_variant(name, None) -> name`: re-stated here).  This script exec()s those methods with `ControlNetBlock`
(nothing is copied into the repo) on a stub `self` at the scaled-down SD1.5 residual pattern, with stub nets that
record what the method asks of them (sub-directory per distinct `save_pattern` index, VAE detached while saving,
:262-281).  The top-level safetensors file it writes is committed as `checkpoint_dir_golden.safetensors` and the
recorded calls as `checkpoint_dir_golden.json`; `tests/test_host_cpu.py` checks that the mirror's `from_pretrained`
reads that file and that its `save_pretrained` writes the same keys, tensors and sub-directories.

    python tests/golden/make_golden_checkpoint_dir.py
"""
import ast
import json
import os
import shutil
import sys
import tempfile
import textwrap
import types
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import safetensors
import safetensors.torch
import torch
from torch import nn

REF = "/root/reference/model/edgestyle_multicontrolnet.py"
HERE = os.path.dirname(os.path.abspath(__file__))
CH = [64, 64, 64, 64, 128, 128, 128, 128, 128, 128, 128, 128]   # tests/test_host_cpu.py TINY: block_out_channels (64, 128, 128, 128)
HW = [8, 8, 8, 4, 4, 4, 2, 2, 2, 1, 1, 1]
PATTERN = [0, None, 1, None, 1, None]


def main():
    src = open(REF).read()
    ns = {"torch": torch, "nn": nn, "os": os, "safetensors": safetensors, "List": List, "Tuple": Tuple, "Union": Union,
          "Optional": Optional, "Dict": Dict, "Any": Any, "Callable": Callable,
          "SAFETENSORS_WEIGHTS_NAME": "diffusion_pytorch_model.safetensors", "WEIGHTS_NAME": "diffusion_pytorch_model.bin",
          "_add_variant": lambda name, variant=None: name, "ControlNetOutput": object}
    methods = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "ControlNetBlock":
            exec(compile(ast.get_source_segment(src, node), REF, "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "EdgeStyleMultiControlNetModel":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in ("state_dict", "load_state_dict", "save_pretrained"):
                    scope = dict(ns)
                    exec(compile(textwrap.dedent(ast.get_source_segment(src, sub)), REF, "exec"), scope)
                    methods[sub.name] = scope[sub.name]
    Block = ns["ControlNetBlock"]
    torch.manual_seed(31)
    g = torch.Generator().manual_seed(32)
    blocks = [Block(c, (s, s), 6) for c, s in zip(CH, HW)] + [Block(CH[-1], (1, 1), 6)]
    with torch.no_grad():
        for b in blocks:
            for p in b.parameters():
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    calls = []

    class StubNet:
        def __init__(self, name, uses_vae):
            self.name = name
            self.config = types.SimpleNamespace(uses_vae=uses_vae)
            self.controlnet_cond_embedding = types.SimpleNamespace(autoencoder="VAE" if uses_vae else None)

        def set_autoencoder(self, vae):
            calls.append(["set_autoencoder", self.name, vae])
            self.controlnet_cond_embedding.autoencoder = vae

        def save_pretrained(self, path, **kw):
            calls.append(["save_pretrained", self.name, os.path.basename(path),
                          self.controlnet_cond_embedding.autoencoder])

    agn, clo, pose = StubNet("agnostic", True), StubNet("clothes", True), StubNet("openpose", False)
    self = types.SimpleNamespace(nets=[agn, pose, clo, pose, clo, pose], multi_controlnet_down_blocks=nn.ModuleList(blocks[:-1]),
                                 multi_controlnet_mid_block=blocks[-1])
    self.state_dict = lambda *a, **k: methods["state_dict"](self, *a, **k)
    tmp = tempfile.mkdtemp()
    try:
        methods["save_pretrained"](self, tmp, save_pattern=PATTERN)
        listing = sorted(os.listdir(tmp))
        shutil.copy(os.path.join(tmp, "diffusion_pytorch_model.safetensors"),
                    os.path.join(HERE, "checkpoint_dir_golden.safetensors"))
    finally:
        shutil.rmtree(tmp)
    # load_state_dict of the reference accepts what its state_dict produced
    sd = methods["state_dict"](self)
    methods["load_state_dict"](self, sd)
    json.dump({"listing": listing, "calls": calls, "pattern": PATTERN, "ch": CH, "hw": HW, "keys": sorted(sd)},
              open(os.path.join(HERE, "checkpoint_dir_golden.json"), "w"), indent=1)
    print("wrote", listing, len(sd), "tensors;", calls)


if __name__ == "__main__":
    sys.exit(main())
