"""CUDA-event timing (20 back-to-back launches, warm L2) of the persistent GEMM against the one-tile kernels on step shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402
dev = "cuda"
ops.set_gemm_workspace(256 << 20)
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
cases = [("gemm", 32768, 320, 320, None, False, True), ("gemm", 32768, 960, 320, None, False, False), ("gemm", 32768, 2560, 320, None, True, False),
         ("gemm", 32768, 320, 1280, None, False, True), ("gemm", 8192, 1920, 640, None, False, False), ("gemm", 8192, 5120, 640, None, True, False),
         ("gemm", 2048, 1280, 5120, None, False, True), ("conv", 32768, 320, 320, (64, 64, 8), False, False), ("conv", 8192, 640, 640, (32, 32, 8), False, False),
         ("conv", 8192, 320, 640, (64, 64, 2), False, False), ("conv", 2048, 1280, 1280, (16, 16, 8), False, False)]
for kind, M, N, K, whn, geglu, res in cases:
    taps = 9 if kind == "conv" else 1
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    b = torch.randn(N, K * taps, device=dev, dtype=torch.float16) * (K * taps) ** -0.5
    out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
    r = torch.randn(M, N, device=dev, dtype=torch.float16) if res else None
    bias = torch.zeros(N, device=dev)
    fl = 2.0 * M * N * K * taps
    row = []
    for bn in ([160, 320, 1160] if geglu else [128, 160, 256, 320, 1128, 1160, 1256]):
        kw = dict(out=out, bias=bias, residual=r, block_n=bn, split_k=1, act=1 if geglu else 0)
        if kind == "conv": kw.update(taps=9, whn=whn, c1=K)
        try:
            t = timeit(lambda: ops.gemm(a, b, N, **kw))
            row.append(f"{bn}:{t:.1f}us/{fl / t / 1e6:.0f}TF")
        except Exception as e:
            row.append(f"{bn}:fail")
    print(f"{kind} M={M} N={N} K={K * taps}{' geglu' if geglu else ''}{' +res' if res else ''}: " + "  ".join(row), flush=True)
