"""BASELINE config 5 (768x1024 image = 96x128 latent, full SD1.5 widths, rank-32 LoRA) on one B200: per-step parity of
the CUDA engine against the fp32 oracle (run on the same GPU) for fp16 and bf16 storage, 20-step DDIM latent PSNR, and
the step time.  Prints one JSON line per dtype; `ES_TUNE_DUMP=path` saves the tile table the eager warm-up measured.

    python tools/config5_probe.py [--hw 96 128] [--steps 20] [--dtypes fp16 bf16] [--no-psnr]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, nargs=2, default=[96, 128])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--dtypes", nargs="+", default=["fp16", "bf16"])
    ap.add_argument("--no-psnr", action="store_true")
    a = ap.parse_args()
    from edgestyle_b200.engine import DenoiseEngine
    from oracle.schedulers import DDIMScheduler
    from oracle.sd15 import SD15Config
    from oracle.step import build_models, denoise, fused_step, synthetic_inputs

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    h, w = a.hw
    cfg = SD15Config()
    m = build_models(cfg, (h, w), rank=32)
    inp = synthetic_inputs(cfg, 1, h, w)
    sds = (m.unet.state_dict(), [m.lora_agnostic.state_dict(), m.lora_clothes.state_dict()], m.openpose.state_dict(),
           m.controlnet.merge_state_dict())
    m.unet.to(dev)
    m.controlnet.to(dev)
    inp.latents, inp.prompt_embeds = inp.latents.to(dev), inp.prompt_embeds.to(dev)
    inp.conds = [c.to(dev) for c in inp.conds]
    x = torch.cat([inp.latents] * 2)
    wants = {t: fused_step(m, x, torch.tensor(t, device=dev), inp.prompt_embeds, inp.conditioning_scale, inp.conds)
             for t in (951, 501, 1)}
    want_lat = None if a.no_psnr else denoise(m, inp, a.steps, 4.5)
    sch = DDIMScheduler()
    ts = sch.set_timesteps(a.steps)
    for name in a.dtypes:
        dtype = torch.float16 if name == "fp16" else torch.bfloat16
        eng = DenoiseEngine(cfg, *sds, rows=2, h=h, w=w, dtype=dtype, use_graph=True)
        eng.set_prompt(inp.prompt_embeds)
        eng.set_conditioning(inp.conds)
        rec = {"config": f"BASELINE configs[4]: {8 * h}x{8 * w} ({h}x{w} latent), CFG batch 2, {name} storage",
               "per_step": {}}
        for t, want in wants.items():
            got = eng.step(x, torch.tensor(t, device=dev), inp.conditioning_scale).float()
            cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
            rec["per_step"][str(t)] = {"cos": round(cos, 6), "max_abs": round((got - want).abs().max().item(), 5),
                                       "rel_rms": round(((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item(), 5),
                                       "eps_absmax": round(want.abs().max().item(), 3)}
        if want_lat is not None:
            lat = inp.latents.clone().float()
            for t in ts:
                eng.step(torch.cat([lat] * 2), torch.tensor(float(t), device=dev), inp.conditioning_scale)
                a_t, a_p = sch.coefficients(int(t))
                eng.cfg_ddim_update(lat, float(a_t), float(a_p), 4.5)
            mse = (lat - want_lat).pow(2).mean().item()
            rec["ddim_psnr_db"] = round(10 * math.log10(want_lat.abs().max().item() ** 2 / max(mse, 1e-30)), 2)
        # step time (graph replay + cfg_ddim), 20 steps
        lat = inp.latents.clone().float()
        x2 = torch.empty_like(x)
        tdev = torch.tensor(501.0, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            e0.record()
            for _ in range(a.steps):
                torch.cat([lat, lat], out=x2)
                eng.step(x2, tdev, inp.conditioning_scale)
                eng.cfg_ddim_update(lat, 0.5, 0.6, 4.5)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        rec["ms_per_step"] = round(ms, 3)
        rec["steps_per_s"] = round(1e3 / ms, 2)
        rec["launches_per_step"] = eng.launches_per_step
        print(json.dumps(rec), flush=True)
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
