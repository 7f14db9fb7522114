"""Sweep block_n x split_k for the small-M (8x8 / 16x16-level) GEMM shapes, cold L2, CUDA events."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ops.set_gemm_workspace(512 << 20)

def timeit(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

for (imgs, hw, cin, cout, taps) in [(8, 8, 1280, 1280, 9), (6, 8, 1280, 1280, 9), (2, 8, 2560, 1280, 9), (8, 8, 1280, 1280, 1),
                                    (2, 8, 1280, 1280, 1), (8, 8, 5120, 1280, 1), (2, 16, 2560, 1280, 9), (8, 16, 1280, 1280, 9)]:
    M = imgs * hw * hw
    x = torch.randn(M, cin, device=dev, dtype=torch.float16)
    wt = torch.randn(cout, taps * cin, device=dev, dtype=torch.float16) * (taps * cin) ** -0.5
    out = torch.empty(M, cout, device=dev, dtype=torch.float16)
    bias = torch.zeros(cout, device=dev)
    res = []
    for bn in (0, 32, 64, 128, 160, 256):
        for sk in (0, 1, 2, 3, 4, 6, 8, 12, 16):
            def run():
                ops.gemm(x, wt, cout, out=out, taps=taps, whn=(hw, hw, imgs) if taps == 9 else None, bias=bias,
                         c1=cin, block_n=bn, split_k=sk)
            try:
                run(); torch.cuda.synchronize()
            except Exception:
                continue
            res.append((timeit(run), bn, sk))
    res.sort()
    auto = [r for r in res if r[1] == 0 and r[2] == 0][0][0]
    print(f"M={M} K={taps * cin} N={cout} taps={taps}: auto {auto:.1f} us; best " +
          ", ".join(f"{t:.1f}us(bn={bn},sk={sk})" for t, bn, sk in res[:5]), flush=True)
