import os, sys, torch
sys.path.insert(0, "/root/repo")
from edgestyle_b200 import ops
ops.set_gemm_workspace(256 << 20)
dev = "cuda"
which = sys.argv[1]
if which == "persist":
    M, N, K = 32768, 320, 320
    x = torch.randn(M, K, device=dev, dtype=torch.float16); wt = torch.randn(N, K, device=dev, dtype=torch.float16) * K ** -0.5
    o = torch.empty(M, N, device=dev, dtype=torch.float16); bias = torch.zeros(N, device=dev); r = torch.randn(M, N, device=dev, dtype=torch.float16)
    for _ in range(3): ops.gemm(x, wt, N, out=o, bias=bias, residual=r, block_n=1160)
else:
    imgs, h, cin, cout = 8, 64, 320, 320
    x = torch.randn(imgs * h * h, cin, device=dev, dtype=torch.float16); wt = torch.randn(cout, 9 * cin, device=dev, dtype=torch.float16) * (9 * cin) ** -0.5
    o = torch.empty(imgs * h * h, cout, device=dev, dtype=torch.float16); bias = torch.zeros(cout, device=dev)
    for _ in range(3): ops.gemm(x, wt, cout, out=o, taps=9, whn=(h, h, imgs), bias=bias, c1=cin, block_n=320)
torch.cuda.synchronize()
