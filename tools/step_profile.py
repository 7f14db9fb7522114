"""Per-op timing of one eager denoise step (CUDA events around every native call) with FLOP / byte accounting.
    python tools/step_profile.py [images] > gpurun_out/step_profile.txt"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import config as C, ops  # noqa: E402
from edgestyle_b200.engine import DenoiseEngine  # noqa: E402
from edgestyle_b200.synth import synth_state_dicts  # noqa: E402

images = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = C.UNetConfig()
h = w = 64
sds = synth_state_dicts(cfg, h, w, rank=32, seed=0)
eng = DenoiseEngine(cfg, sds["unet"], sds["lora"], sds["pose"], sds["merge"], rows=2 * images, h=h, w=w)
g = torch.Generator().manual_seed(1)
eng.set_prompt(torch.randn(2 * images, 77, 768, generator=g))
eng.set_conditioning([torch.randn(2 * images, 320, h, w, generator=g) * 0.5 for _ in range(6)])
x = torch.randn(2 * images, 4, h, w, generator=g).cuda()
records = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def wrap(name, fn, describe):
    def inner(*a, **k):
        desc, flops, nbytes = describe(*a, **k)
        if os.environ.get("COLD"):
            flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.LAUNCHES
        e0.record()
        r = fn(*a, **k)
        e1.record()
        torch.cuda.synchronize()
        records.append((name, desc, e0.elapsed_time(e1) * 1e3, flops, nbytes, ops.LAUNCHES - n0))
        return r
    return inner


def d_gemm(a, b, n, **k):
    M = a.shape[0]
    taps = k.get("taps", 1)
    c1 = k.get("c1") or a.shape[1]
    K = taps * c1 + (k["a2"].shape[1] if k.get("a2") is not None else 0)
    kind = "conv3x3" if taps == 9 else "gemm"
    extra = ("+lora" if k.get("a2") is not None and taps == 1 else "") + ("+sc" if k.get("a2") is not None and taps == 9 else "") \
        + ("+geglu" if k.get("act") else "")
    return f"{kind}{extra} M={M} N={n} K={K}", 2.0 * M * n * K, 2.0 * (M * K / (9 if taps == 9 else 1) + n * K + M * n)


def d_att(q, kk, v, out, batch, heads, nq, nkv, scale=None):
    d = q.shape[1] // heads
    return f"attn b={batch} nq={nq} nkv={nkv} d={d}", 4.0 * batch * heads * nq * nkv * d, 2.0 * (2 * batch * nq + 2 * batch * nkv) * heads * d


def d_gn(x0, out, *a, **k):
    return f"groupnorm M={x0.shape[0]} C={out.shape[1]}", 0.0, 6.0 * out.numel()


def d_ln(x, out, *a, **k):
    return f"layernorm M={x.shape[0]} C={x.shape[1]}", 0.0, 4.0 * x.numel()


def d_merge(levels, scale_dev, B):
    # bytes: phase 1 reads 6 residual slabs, phase 2 reads them again + g1/be1 and writes z, phase 3 reads z, g2/be2, skip
    # and writes dst (16-bit everything except an fp32 z)
    tot = 0.0
    for lv in levels:
        n = B * lv["hw"] * lv["C"]
        zb = lv["z"].element_size()
        tot += n * (12 + 12 + zb + zb + 2 + 2) + lv["hw"] * lv["C"] * (12 + 4)
    return f"merge B={B} levels={len(levels)}", 0.0, tot


def d_other(*a, **k):
    return "", 0.0, 0.0


ops.gemm = wrap("gemm", ops.gemm, d_gemm)
ops.attention = wrap("attention", ops.attention, d_att)
ops.groupnorm = wrap("groupnorm", ops.groupnorm, d_gn)
ops.layernorm = wrap("layernorm", ops.layernorm, d_ln)
ops.merge_levels = wrap("merge", ops.merge_levels, d_merge)
for nm in ("small_linear", "im2col3x3", "upsample2x", "nchw_to_nhwc", "timestep_embedding"):
    setattr(ops, nm, wrap(nm, getattr(ops, nm), d_other))

eng.step(x, 500.0)  # warm-up (allocations)
records.clear()
eng.step(x, 500.0)
import json
json.dump([(r[0], r[1], r[3], r[4], r[5]) for r in records], open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "oplog.json"), "w"))
tot = sum(r[2] for r in records)
print(f"ops {len(records)}  total {tot / 1e3:.2f} ms  ({'cold' if os.environ.get('COLD') else 'warm'} L2)")
by = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for name, desc, us, fl, by_, _nl in records:
    k = (name, desc)
    by[k][0] += 1
    by[k][1] += us
    by[k][2] += fl
    by[k][3] += by_
fam = collections.defaultdict(float)
for (name, desc), v in by.items():
    fam[name] += v[1]
print({k: round(v / 1e3, 2) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
print(f"{'us total':>9s} {'n':>3s} {'us avg':>8s} {'TFLOP/s':>8s} {'GB/s':>7s}  op")
for (name, desc), (n, us, fl, by_) in sorted(by.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{us:9.1f} {n:3d} {us / n:8.1f} {fl / us / 1e6:8.1f} {by_ / us / 1e3:7.0f}  {name} {desc}")
