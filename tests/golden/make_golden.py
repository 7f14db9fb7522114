"""Generate golden vectors from the reference's OWN code (run in the build container only).

The three hot-path files of the reference import diffusers at module scope and cannot be imported
(SURVEY.md F3), but `ControlNetBlock`, `interleave_tensors*`, `zero_module`, `ones_module`
(/root/reference/model/edgestyle_multicontrolnet.py:23-63,467-514) are pure torch.  This script
exec()s exactly those source spans (nothing is copied into the repo) on seeded inputs and stores
inputs + parameters + outputs in `merge_block_golden.pt`, which `tests/test_oracle_golden.py`
replays against `oracle.merge`.

    python tests/golden/make_golden.py
"""
import ast
import os
import sys

import torch
from torch import nn
from typing import List, Tuple, Union, Optional, Dict, Any, Callable

REF = "/root/reference/model/edgestyle_multicontrolnet.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "merge_block_golden.pt")


def load_reference_symbols():
    src = open(REF).read()
    tree = ast.parse(src)
    wanted = {"ControlNetBlock", "interleave_tensors", "interleave_tensors_from_list_of_lists", "zero_module",
              "ones_module"}
    ns = {"torch": torch, "nn": nn, "List": List, "Tuple": Tuple, "Union": Union, "Optional": Optional,
          "Dict": Dict, "Any": Any, "Callable": Callable, "ControlNetOutput": object}
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in wanted:
            code = ast.get_source_segment(src, node)
            exec(compile(code, REF, "exec"), ns)
    return ns


def main():
    ns = load_reference_symbols()
    Block, interleave = ns["ControlNetBlock"], ns["interleave_tensors"]
    cases = []
    g = torch.Generator().manual_seed(20240607)
    for (c, h, w, b) in [(8, 4, 4, 2), (16, 8, 4, 3), (32, 2, 2, 1)]:
        torch.manual_seed(c * 1000 + h)
        blk = Block(c, (h, w), 6)
        with torch.no_grad():
            for p in blk.parameters():  # move LN affine away from (1, 0) so every term matters
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
        res = [torch.randn(b, c, h, w, generator=g) for _ in range(6)]
        with torch.no_grad():
            inter = interleave(res)
            out = blk(inter)
        cases.append({"shape": (c, h, w, b), "state_dict": {k: v.clone() for k, v in blk.state_dict().items()},
                      "residuals": res, "interleaved": inter, "out": out})
    torch.save(cases, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
