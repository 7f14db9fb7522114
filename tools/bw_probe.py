"""Bandwidth probe of the norm / merge kernels vs a plain torch copy of the same bytes (CUDA events, cold L2)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops  # noqa: E402
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def t(fn, reps=10, cold=True):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

for rows, c, imgs in [(32768, 320, 8), (8192, 640, 8), (2048, 1280, 8), (512, 1280, 8), (8192, 320, 2)]:
    x = torch.randn(rows, c, device=dev, dtype=torch.float16)
    y = torch.empty_like(x)
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev)
    ws = torch.zeros(imgs, 32, 2, device=dev)
    nbytes = x.numel() * 2
    for cold in (True, False):
        tc = t(lambda: y.copy_(x), cold=cold)
        tl = t(lambda: ops.layernorm(x, y, g, b), cold=cold)
        tg = t(lambda: ops.groupnorm(x, y, g, b, ws, imgs, rows // imgs, 32, 1e-5, True), cold=cold)
        print(f"[{rows}x{c}] {'cold' if cold else 'warm'}: copy {tc:6.1f} us ({2*nbytes/tc/1e3:5.0f} GB/s)  layernorm {tl:6.1f} us ({2*nbytes/tl/1e3:5.0f} GB/s)"
              f"  groupnorm(memset+stats+apply) {tg:6.1f} us ({3*nbytes/tg/1e3:5.0f} GB/s)", flush=True)

# GroupNorm apply alone (statistics already accumulated by the producing GEMM's epilogue: the common case in the step)
for rows, c, imgs in [(32768, 320, 8), (8192, 640, 8), (262144, 128, 1)]:
    x = torch.randn(rows, c, device=dev, dtype=torch.float16)
    y = torch.empty_like(x)
    g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev)
    ws = torch.zeros(imgs, 32, 2, device=dev)
    ops.groupnorm(x, y, g, b, ws, imgs, rows // imgs, 32, 1e-5, True)  # leaves valid statistics in ws
    nbytes = x.numel() * 2
    for cold in (True, False):
        tc = t(lambda: y.copy_(x), cold=cold)
        ta = t(lambda: ops.groupnorm(x, y, g, b, ws, imgs, rows // imgs, 32, 1e-5, True, stats_ready=True), cold=cold)
        print(f"[{rows}x{c}] {'cold' if cold else 'warm'}: copy {tc:6.1f} us ({2*nbytes/tc/1e3:5.0f} GB/s)  gn_apply {ta:6.1f} us "
              f"({2*nbytes/ta/1e3:5.0f} GB/s)", flush=True)

# row softmax of the VAE mid-block attention (fp32 scores in, fp16 probabilities out): 6 B per score
for n in (1024, 4096):
    s = torch.randn(n, n, device=dev)
    p = torch.empty(n, n, device=dev, dtype=torch.float16)
    nb = s.numel() * 6
    for cold in (True, False):
        tc = t(lambda: p.copy_(s), cold=cold)
        ts_ = t(lambda: ops.softmax_rows(s, p), cold=cold)
        print(f"[softmax {n}x{n}] {'cold' if cold else 'warm'}: fp32->fp16 copy {tc:6.1f} us ({nb/tc/1e3:5.0f} GB/s)  softmax_rows {ts_:6.1f} us "
              f"({nb/ts_/1e3:5.0f} GB/s)", flush=True)
