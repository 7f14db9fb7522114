"""Does a K-block-major weight layout (contiguous B tiles) speed up weight-streaming (small-M) GEMMs?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ops
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
    return sorted(ts)[len(ts) // 2]

def block(w, taps, c1):  # [n, taps*c1] -> [taps*kb][n][64]
    n = w.shape[0]
    kb = (c1 + 63) // 64
    wp = torch.zeros(n, taps, kb * 64, device=w.device, dtype=w.dtype)
    wp[:, :, :c1] = w.view(n, taps, c1)
    return wp.view(n, taps * kb, 64).permute(1, 0, 2).contiguous()

for (imgs, hw, cin, cout, taps, bn, sk) in [(8, 8, 1280, 1280, 9, 64, 3), (2, 16, 2560, 1280, 9, 128, 6), (2, 8, 2560, 1280, 9, 64, 12),
                                            (8, 8, 1280, 1280, 1, 32, 1), (8, 16, 1280, 1280, 9, 160, 1), (8, 64, 320, 320, 9, 160, 1)]:
    M = imgs * hw * hw
    x = torch.randn(M, cin, device=dev, dtype=torch.float16)
    wt = torch.randn(cout, taps * cin, device=dev, dtype=torch.float16) * (taps * cin) ** -0.5
    wb = block(wt, taps, cin)
    out = torch.empty(M, cout, device=dev, dtype=torch.float16)
    out2 = torch.empty(M, cout, device=dev, dtype=torch.float16)
    bias = torch.zeros(cout, device=dev)
    whn = (hw, hw, imgs) if taps == 9 else None
    f1 = lambda: ops.gemm(x, wt, cout, out=out, taps=taps, whn=whn, bias=bias, c1=cin, block_n=bn, split_k=sk)
    f2 = lambda: ops.gemm(x, wb, cout, out=out2, taps=taps, whn=whn, bias=bias, c1=cin, block_n=bn, split_k=sk, b_blocked=True)
    f1(); f2(); torch.cuda.synchronize()
    print(f"M={M} K={taps * cin} N={cout} bn={bn} sk={sk}: row-major {timeit(f1):.1f} us, K-block-major {timeit(f2):.1f} us, "
          f"max diff {(out.float() - out2.float()).abs().max().item():.2e}", flush=True)
