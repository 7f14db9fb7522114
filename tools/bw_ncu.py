"""One launch of each bandwidth kernel at step / VAE shapes, for an `ncu --set full` capture (DRAM bytes, achieved
DRAM throughput).  ncu flushes the caches before every kernel, so these are cold-L2 numbers.

    ncu --set full --clock-control none -k regex:"gn_|softmax|layernorm" -o prof python tools/bw_ncu.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402

dev = "cuda"
for rows, c, imgs in [(32768, 320, 8), (8192, 640, 8), (262144, 128, 1), (196608, 256, 3)]:
    x = torch.randn(rows, c, device=dev, dtype=torch.float16)
    y = torch.empty_like(x)
    g = torch.ones(c, device=dev)
    b = torch.zeros(c, device=dev)
    ws = torch.zeros(imgs, 32, 2, device=dev)
    ops.groupnorm(x, y, g, b, ws, imgs, rows // imgs, 32, 1e-5, True)   # gn_stats + gn_apply
    if c % 8 == 0 and c <= 2048 and rows <= 32768:
        ops.layernorm(x, y, g, b)
s = torch.randn(4096, 4096, device=dev)
p = torch.empty(4096, 4096, device=dev, dtype=torch.float16)
ops.softmax_rows(s, p)
torch.cuda.synchronize()
print("done")
