"""GPU *library* baseline of the denoise step: the fp32 oracle's modules cast to fp16 channels-last and run by
PyTorch's stock dispatch (cuDNN convolutions, cuBLAS GEMMs, SDPA flash attention, ATen norm / elementwise kernels)
on the same B200, eager and captured into one CUDA graph.  This is the bar SURVEY.md 2.1 / BASELINE.md 3 set: the
fused single-step form of /root/reference/export_onnx.py:43-74 as the reference's own stack would execute it on this
GPU (diffusers itself is absent offline, so the restated modules stand in for it -- same operators, same shapes).

    python tools/lib_baseline.py [--images 1] [--hw 64 64] [--steps 20] [--dtype fp16|bf16]

Baseline infrastructure like bench.py's cpu_baseline leg: never on the product path.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def gpu_library_baseline(images: int = 1, h: int = 64, w: int = 64, steps: int = 20, warmup: int = 3,
                         dtype=torch.float16, models=None, device="cuda", guidance: float = 4.5) -> dict:
    """Denoise steps/s of the library path (UNet + 6 ControlNet branches + merge + CFG + DDIM), eager and graphed."""
    from oracle.schedulers import DDIMScheduler
    from oracle.sd15 import SD15Config
    from oracle.step import build_models, cfg_combine, fused_step, synthetic_inputs

    cfg = SD15Config()
    t_build = time.perf_counter()
    m = models if models is not None else build_models(cfg, (h, w), rank=32)
    m.unet.to(device=device, dtype=dtype).to(memory_format=torch.channels_last)
    m.controlnet.to(device=device, dtype=dtype).to(memory_format=torch.channels_last)
    t_build = time.perf_counter() - t_build
    inp = synthetic_inputs(cfg, images, h, w)
    pe = inp.prompt_embeds.to(device=device, dtype=dtype)
    conds = [c.to(device=device, dtype=dtype).contiguous(memory_format=torch.channels_last) for c in inp.conds]
    lat = inp.latents.to(device=device, dtype=torch.float32)
    sch = DDIMScheduler()
    ts = [torch.tensor(float(t), device=device) for t in sch.set_timesteps(20)]
    coefs = [sch.coefficients(int(t)) for t in sch.set_timesteps(20)]
    scale = [1.0] * 6
    x_static = torch.zeros(2 * images, cfg.in_channels, h, w, device=device, dtype=dtype).contiguous(
        memory_format=torch.channels_last)
    t_static = torch.zeros((), device=device)

    def body():
        return fused_step(m, x_static, t_static, pe, scale, conds)

    def update(eps, k):
        a_t, a_p = coefs[k]
        e = cfg_combine(eps.float(), guidance)
        x0 = (lat - math.sqrt(1 - a_t) * e) / math.sqrt(a_t)
        lat.copy_(math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * e)

    def step_eager(i):
        k = i % 20
        x_static.copy_(torch.cat([lat, lat]).to(dtype))
        t_static.copy_(ts[k])
        update(body(), k)

    out = {"dtype": str(dtype).split(".")[-1], "images": images, "latent": [h, w], "steps": steps,
           "what": "oracle modules in 16-bit channels-last through stock PyTorch dispatch (cuDNN / cuBLAS / SDPA / ATen)",
           "build_s": round(t_build, 1)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        torch.backends.cudnn.benchmark = True
        for i in range(warmup):
            step_eager(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            step_eager(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["eager"] = {"steps_per_s": round(1e3 / ms, 3), "ms_per_step": round(ms, 3)}
        # ---- one CUDA graph per step body (the scheduler update stays eager, as in our own arm) ----
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eps_static = body()

            def step_graph(i):
                k = i % 20
                x_static.copy_(torch.cat([lat, lat]).to(dtype))
                t_static.copy_(ts[k])
                g.replay()
                update(eps_static, k)

            lat.copy_(inp.latents.to(device))
            for i in range(warmup):
                step_graph(i)
            torch.cuda.synchronize()
            e0.record()
            for i in range(steps):
                step_graph(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out["graphed"] = {"steps_per_s": round(1e3 / ms, 3), "ms_per_step": round(ms, 3)}
        except Exception as exc:  # capture can fail on a host-synchronising op: report, keep the eager number
            out["graphed"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    out["finite"] = bool(torch.isfinite(lat).all().item())
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1)
    ap.add_argument("--hw", type=int, nargs=2, default=[64, 64])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    a = ap.parse_args()
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    print(json.dumps({"gpu_library_baseline": gpu_library_baseline(a.images, a.hw[0], a.hw[1], a.steps, dtype=dt)}),
          flush=True)
