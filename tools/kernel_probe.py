"""Tiny driver for ncu --set full captures and CUDA-event timings of the two hot kernels at step shapes.
    python tools/kernel_probe.py [attn|conv|all] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda"
ops.set_gemm_workspace(256 << 20)


def timeit(fn, name, flops):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name}: {us:.1f} us  {flops / us / 1e6:.1f} TFLOP/s", flush=True)


if what in ("attn", "all"):
    for (batch, heads, d, n) in [(8, 8, 40, 4096), (8, 8, 80, 1024), (8, 8, 160, 256)]:
        C = heads * d
        qkv = torch.randn(batch * n, 3 * C, device=dev, dtype=torch.float16)
        out = torch.empty(batch * n, C, device=dev, dtype=torch.float16)
        timeit(lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, batch, heads, n, n),
               f"attention b{batch} h{heads} d{d} n{n}", 4.0 * batch * heads * n * n * d)
if what in ("pair", "all"):
    # CTA-pair kernel (cta_group::2, 256 x 320 tiles) on the shapes the tuner gives it: GEGLU feed-forward, long-K linear
    for (M, N, K, geglu) in [(32768, 2560, 320, True), (32768, 320, 1280, False)]:
        x = torch.randn(M, K, device=dev, dtype=torch.float16)
        wt = torch.randn(N, K, device=dev, dtype=torch.float16) * K ** -0.5
        o = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
        bias = torch.zeros(N, device=dev)
        timeit(lambda: ops.gemm(x, wt, N, out=o, bias=bias, act=1 if geglu else 0, block_n=320),
               f"pair gemm M={M} N={N} K={K}{' geglu' if geglu else ''}", 2.0 * M * N * K)
if what in ("conv", "all"):
    for (imgs, h, w, cin, cout) in [(8, 64, 64, 320, 320), (8, 32, 32, 640, 640), (8, 8, 8, 1280, 1280)]:
        x = torch.randn(imgs * h * w, cin, device=dev, dtype=torch.float16)
        wt = torch.randn(cout, 9 * cin, device=dev, dtype=torch.float16) * (9 * cin) ** -0.5
        o = torch.empty(imgs * h * w, cout, device=dev, dtype=torch.float16)
        bias = torch.zeros(cout, device=dev)
        timeit(lambda: ops.gemm(x, wt, cout, out=o, taps=9, whn=(w, h, imgs), bias=bias, c1=cin),
               f"conv3x3 {imgs}x{h}x{w} {cin}->{cout}", 2.0 * imgs * h * w * cout * 9 * cin)
