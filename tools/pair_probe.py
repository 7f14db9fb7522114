"""Cold-L2 timing of the CTA-pair kernel (block_n 320) against the single-CTA tile widths on the step's big shapes.
    python tools/pair_probe.py > gpurun_out/pair_probe.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ops.set_gemm_workspace(512 << 20)


def time_call(fn, reps=7):
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


cases = []
for imgs, hw, cin, cout in [(8, 64, 320, 320), (6, 64, 320, 320), (2, 64, 960, 320), (2, 64, 640, 320), (8, 32, 640, 640), (8, 32, 320, 640),
                            (2, 32, 1280, 640), (8, 16, 1280, 1280), (2, 16, 2560, 1280), (8, 8, 1280, 1280), (2, 8, 2560, 1280)]:
    cases.append(("conv", imgs * hw * hw, cout, cin, (hw, hw, imgs), False, False))
for M, N, K, geglu, res in [(32768, 320, 320, False, True), (32768, 960, 320, False, False), (32768, 2560, 320, True, False),
                            (32768, 320, 1280, False, True), (8192, 640, 640, False, True), (8192, 1920, 640, False, False),
                            (8192, 5120, 640, True, False), (8192, 640, 2560, False, True), (2048, 1280, 1280, False, True),
                            (2048, 10240, 1280, True, False), (2048, 1280, 5120, False, True), (8192, 320, 320, False, True),
                            (8192, 2560, 320, True, False), (2048, 640, 640, False, True), (2048, 5120, 640, True, False),
                            (512, 1280, 1280, False, True), (512, 10240, 1280, True, False), (512, 1280, 5120, False, True)]:
    cases.append(("gemm", M, N, K, None, geglu, res))
print(f"{'case':44s} {'bn':>4s} {'sk':>3s} {'us':>8s} {'TFLOP/s':>8s}")
for kind, M, N, K, whn, geglu, res in cases:
    taps = 9 if kind == "conv" else 1
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    b = torch.randn(N, K * taps, device=dev, dtype=torch.float16) * (K * taps) ** -0.5
    out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
    r = torch.randn(M, N, device=dev, dtype=torch.float16) if res else None
    bias = torch.zeros(N, device=dev)
    flops = 2.0 * M * N * K * taps
    name = f"{kind} M={M} N={N} K={K * taps}{' geglu' if geglu else ''}{' +res' if res else ''}"
    best = None
    for bn in ([160, 320, 1160] if geglu else [64, 128, 160, 256, 320, 1128, 1160, 1256]):
        for sk in ((1,) if bn >= 1000 else (1, 0)):
            kw = dict(out=out, bias=bias, residual=r, block_n=bn, split_k=sk, act=1 if geglu else 0)
            if kind == "conv":
                kw.update(taps=9, whn=whn, c1=K)
            try:
                t = time_call(lambda: ops.gemm(a, b, N, **kw))
            except Exception as e:  # noqa: BLE001
                print(f"{name:44s} {bn:4d} {sk:3d}  failed: {e}")
                continue
            print(f"{name:44s} {bn:4d} {sk:3d} {t:8.1f} {flops / t / 1e6:8.1f}")
            if best is None or t < best[0]:
                best = (t, bn, sk)
    print(f"  -> best {name}: {best[0]:.1f} us bn={best[1]} sk={best[2]}\n")
