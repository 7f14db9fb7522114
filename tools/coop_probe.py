"""Cold-L2 timing of the cooperative split-K (negative split_k) against the last-arriver split-K on the small-M, long-K
convolutions of the 8x8 / 16x16 levels (weight streaming: the HBM floor is weights / 6.5 TB/s)."""
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
for imgs, hw, cin, cout in [(8, 8, 1280, 1280), (6, 8, 1280, 1280), (2, 8, 2560, 1280), (2, 16, 2560, 1280), (2, 16, 1920, 1280), (8, 16, 1280, 1280), (2, 32, 1280, 640)]:
    M = imgs * hw * hw
    a = torch.randn(M, cin, device=dev, dtype=torch.float16)
    b = torch.randn(cout, 9 * cin, device=dev, dtype=torch.float16) * (9 * cin) ** -0.5
    out = torch.empty(M, cout, device=dev, dtype=torch.float16)
    bias = torch.zeros(cout, device=dev)
    wbytes = b.numel() * 2
    res = []
    for bn in (64, 128, 256):
        tiles = ((M + 127) // 128) * ((cout + bn - 1) // bn)
        kb = 9 * cin // 64
        cands = [0] + ([-(148 // tiles), -max(2, 148 // tiles // 2)] if tiles <= 74 else [])
        for sk in cands:
            try:
                t = time_call(lambda: ops.gemm(a, b, cout, out=out, taps=9, whn=(hw, hw, imgs), bias=bias, c1=cin, block_n=bn, split_k=sk))
                res.append((t, bn, sk))
            except Exception as e:
                res.append((1e9, bn, sk))
    res.sort()
    print(f"conv {imgs}x{hw}x{hw} {cin}->{cout} (M={M}, weights {wbytes / 1e6:.0f} MB, HBM floor {wbytes / 6.55e6:.1f} us): "
          + "  ".join(f"bn{bn}/sk{sk}:{t:.1f}us" for t, bn, sk in res[:6]), flush=True)
