"""GPU micro-benchmark: es_gemm over the step's representative layer shapes x (block_n, stages, split_k).
Cold-L2 timing (a 512 MB buffer is rewritten between launches).  Run on the B200 box:
    python tools/gemm_sweep.py > gpurun_out/gemm_sweep.txt
"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ops.set_gemm_workspace(512 << 20)


def time_call(fn, reps=6):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def conv_case(imgs, h, w, cin, cout):
    x = torch.randn(imgs * h * w, cin, device=dev, dtype=torch.float16)
    wt = torch.randn(cout, 9 * cin, device=dev, dtype=torch.float16) * (9 * cin) ** -0.5
    out = torch.empty(imgs * h * w, cout, device=dev, dtype=torch.float16)
    bias = torch.zeros(cout, device=dev)
    flops = 2.0 * imgs * h * w * cout * 9 * cin
    def run(bn, st, sk):
        ops.gemm(x, wt, cout, out=out, taps=9, whn=(w, h, imgs), bias=bias, c1=cin, block_n=bn, stages=st, split_k=sk)
    return f"conv3x3 {imgs}x{h}x{w} {cin}->{cout}", flops, run


def lin_case(M, K, N, geglu=False):
    x = torch.randn(M, K, device=dev, dtype=torch.float16)
    wt = torch.randn(N, K, device=dev, dtype=torch.float16) * K ** -0.5
    out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
    bias = torch.zeros(N, device=dev)
    flops = 2.0 * M * N * K
    def run(bn, st, sk):
        ops.gemm(x, wt, N, out=out, bias=bias, act=1 if geglu else 0, block_n=bn, stages=st, split_k=sk)
    return f"linear M={M} K={K} N={N}{' geglu' if geglu else ''}", flops, run


cases = [
    conv_case(8, 64, 64, 320, 320), conv_case(6, 64, 64, 320, 320), conv_case(2, 64, 64, 960, 320),
    conv_case(8, 32, 32, 640, 640), conv_case(2, 32, 32, 1920, 640),
    conv_case(8, 16, 16, 1280, 1280), conv_case(2, 16, 16, 2560, 1280),
    conv_case(8, 8, 8, 1280, 1280), conv_case(6, 8, 8, 1280, 1280), conv_case(2, 8, 8, 2560, 1280),
    lin_case(32768, 320, 2560, True), lin_case(32768, 1280, 320), lin_case(32768, 320, 960), lin_case(32768, 320, 320),
    lin_case(8192, 640, 5120, True), lin_case(8192, 2560, 640), lin_case(8192, 640, 1920),
    lin_case(2048, 1280, 10240, True), lin_case(2048, 5120, 1280), lin_case(2048, 1280, 3840),
    lin_case(512, 1280, 10240, True), lin_case(512, 5120, 1280), lin_case(512, 11520, 1280), lin_case(128, 1280, 1280),
]
configs = [(0, 0, 1), (0, 0, 0)]
for bn in (64, 128, 160, 256):
    configs.append((bn, 0, 1))
for bn in (64, 128, 256):
    for sk in (2, 4, 8, 16):
        configs.append((bn, 0, sk))
for bn in (128, 256):
    for sk in (2, 4, 8):
        configs.append((bn, 4, sk))

print(f"{'case':44s} {'bn':>4s} {'st':>3s} {'sk':>3s} {'us':>9s} {'TFLOP/s':>8s}")
for name, flops, run in cases:
    best = None
    for bn, st, sk in configs:
        try:
            run(bn, st, sk)
            torch.cuda.synchronize()
        except Exception as e:  # invalid combination for this shape (e.g. GEGLU divisibility)
            continue
        us = time_call(lambda: run(bn, st, sk))
        tf = flops / us / 1e6
        print(f"{name:44s} {bn:4d} {st:3d} {sk:3d} {us:9.1f} {tf:8.1f}")
        if best is None or us < best[0]:
            best = (us, bn, st, sk)
    print(f"  -> best {name}: {best[0]:.1f} us bn={best[1]} stages={best[2]} split_k={best[3]}\n", flush=True)
