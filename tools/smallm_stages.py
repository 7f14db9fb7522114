"""Small-M, long-K convolution (8x8 level: M = 512, N = 1280, K = 11520): time against ring depth / split factor / tile width
(cold L2).  Is the weight stream latency-bound (deeper ring helps) or bandwidth-bound?"""
import os, sys
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
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
imgs, hw, cin, cout = 8, 8, 1280, 1280
M = imgs * hw * hw
a = torch.randn(M, cin, device=dev, dtype=torch.float16)
b = torch.randn(cout, 9 * cin, device=dev, dtype=torch.float16) * (9 * cin) ** -0.5
out = torch.empty(M, cout, device=dev, dtype=torch.float16)
bias = torch.zeros(cout, device=dev)
for bn in (64, 128, 256):
    for sk in (1, 2, 3, 6, 12):
        row = []
        for st in (0, 2, 4, 6, 8, 12):
            try:
                t = time_call(lambda: ops.gemm(a, b, cout, out=out, taps=9, whn=(hw, hw, imgs), bias=bias, c1=cin, block_n=bn, split_k=sk, stages=st))
                row.append(f"st{st}:{t:.1f}")
            except Exception as e:
                row.append(f"st{st}:fail")
        tiles = 4 * ((cout + bn - 1) // bn)
        print(f"bn={bn} sk={sk} ({tiles * sk} CTAs): " + "  ".join(row), flush=True)
