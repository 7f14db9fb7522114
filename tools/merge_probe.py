"""The EdgeStyle merge (13 ControlNetBlocks of one step, B = 2, 64x64 latent) as the engine launches it: one call of
ops.merge_levels per level group = 3 launches each.  CUDA-event timing (cold / warm L2) against the algorithmic bytes
(>= 295.6 MB per step: six residual slabs read twice, g1/be1, z written and read, g2/be2, skip, dst).
    python tools/merge_probe.py            # timing
    ncu --set full --clock-control none -k regex:merge_levels python tools/merge_probe.py once   # DRAM counters
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import config as C, ops  # noqa: E402
from edgestyle_b200.engine import pack_merge_block  # noqa: E402
from edgestyle_b200.synth import synth_state_dicts  # noqa: E402

dev = "cuda"
cfg = C.UNetConfig()
B, h, w = 2, 64, 64
sds = synth_state_dicts(cfg, h, w, rank=4, seed=0)
shapes = C.residual_shapes(cfg, h, w)
pfx = [f"multi_controlnet_down_blocks.{i}." for i in range(len(shapes) - 1)] + ["multi_controlnet_mid_block."]
levels = []
alg = 0.0
for p, (c, hh, ww) in zip(pfx, shapes):
    sub = {k[len(p):]: v for k, v in sds["merge"].items() if k.startswith(p)}
    prm = pack_merge_block(sub, c, hh, ww, torch.float16, dev)
    n = B * hh * ww
    res = [torch.randn(n, c, device=dev, dtype=torch.float16) * 0.1 for _ in range(6)]
    levels.append(dict(res=res, prm=prm, stats=torch.zeros(B, 4, device=dev, dtype=torch.float64),
                       z=torch.empty(n, c, device=dev, dtype=torch.float16), hw=hh * ww, C=c,
                       dst=torch.empty(n, c, device=dev, dtype=torch.float16), skip=torch.randn(n, c, device=dev, dtype=torch.float16)))
    alg += n * c * (12 + 12 + 2 + 2 + 2 + 2) + hh * ww * c * (12 + 4)
scale = torch.ones(6, device=dev)
groups = [levels[3:][::-1], levels[:3][::-1]]  # the engine's default: deep levels first, the three 64x64 levels second


def run():
    for lv in levels:
        lv["stats"].zero_()
    for g in groups:
        ops.merge_levels(g, scale, B)


if len(sys.argv) > 1 and sys.argv[1] == "once":
    run(); torch.cuda.synchronize(); run(); torch.cuda.synchronize(); sys.exit(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for cold in (True, False):
    ts = []
    for _ in range(10):
        for lv in levels:
            lv["stats"].zero_()
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for g in groups:
            ops.merge_levels(g, scale, B)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"merge, 13 levels, 6 launches, {'cold' if cold else 'warm'} L2: {t:.1f} us, algorithmic {alg / 1e6:.1f} MB -> {alg / t / 1e3:.0f} GB/s "
          f"({100 * alg / t / 1e3 / 6554.6:.1f} % of the measured 6.55 TB/s copy peak)")
