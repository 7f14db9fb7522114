#!/usr/bin/env python
"""bench.py -- EdgeStyle denoise hot path on B200: denoise steps/s at 512x512 (64x64 latent).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--images I]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = one full denoise step of BASELINE config 2 on every rank: SD1.5 UNet + six ControlNet/ControlLoRA
branches + EdgeStyle merge + CFG combine + DDIM update over a CFG pair (2 rows) per image, `--images` images per
GPU (default 1), fp16 storage / fp32 accumulate, synthetic seeded weights and inputs.  `value` = denoise steps/s
summed over all ranks (weak scaling: one replica and its own images per GPU, final latents gathered with NCCL).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "denoise_steps_per_s"
UNIT = "steps/s"
# algorithmic FLOPs per step at CFG batch 2, 64x64 latent (UNet 401.64 + 6 x 134.28 GMAC per row =
# 4.829 TFLOP) MINUS the step-invariant text K/V projections (5.56 GMAC/row = 0.022 TFLOP) that the engine computes
# once per prompt instead of every step -- only executed work is credited.
TFLOP_PER_STEP_B2 = 4.807
# DRAM bytes per step of the dominant kernel family + the rest of the step, from profiles/r1c_launches_final.csv
# (ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 673 launches of one step, caches flushed per kernel)
DRAM_BYTES_PER_STEP_B2 = 8.32e9


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1413.1), d.get("hbm_gbs", 6554.6), "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_ours(images: int, device, seed: int):
    from edgestyle_b200 import config as C
    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      EdgeStyleStableDiffusionControlNetPipeline, UNet2DConditionModel)
    from edgestyle_b200.synth import synth_state_dicts

    cfg = C.UNetConfig()
    h = w = 64
    sds = synth_state_dicts(cfg, h, w, rank=32, seed=seed, device="cpu")
    unet = UNet2DConditionModel(cfg, sds["unet"])
    agn = ControlLoRAModel(cfg, sds["lora"][0], lora_linear_rank=32, unet=unet)
    clo = ControlLoRAModel(cfg, sds["lora"][1], lora_linear_rank=32, unet=unet)
    pose = CachedControlNetModel(cfg, sds["pose"])
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], sds["merge"], (h, w))
    pipe = EdgeStyleStableDiffusionControlNetPipeline(unet=unet, controlnet=multi, use_graph=True)
    g = torch.Generator().manual_seed(1234 + seed)
    host = {
        "latents": torch.randn(images, 4, h, w, generator=g).pin_memory(),
        "prompt_embeds": torch.randn(images, 77, 768, generator=g).pin_memory(),
        "negative_prompt_embeds": torch.randn(images, 77, 768, generator=g).pin_memory(),
        "conds": [(torch.randn(2 * images, 320, h, w, generator=g) * 0.5).pin_memory() for _ in range(6)],
    }
    return cfg, pipe, multi, host, (h, w)


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from edgestyle_b200 import build, ext, ops

    if not os.path.exists(ext.LIB_PATH):
        build.build()
    images = args.images
    cfg, pipe, multi, host, (h, w) = build_ours(images, dev, seed=rank)
    B = 2 * images
    eng = multi.engine(B, h, w, use_graph=True)
    from edgestyle_b200.schedulers import DDIMScheduler

    sch = DDIMScheduler()
    ts = sch.set_timesteps(20)
    # ---------------- device-resident timed region: K steps of (graph replay + CFG/DDIM) ----------------
    eng.set_prompt(torch.cat([host["negative_prompt_embeds"], host["prompt_embeds"]]).to(dev))
    eng.set_conditioning([c.to(dev) for c in host["conds"]])
    lat = host["latents"].to(dev).clone()
    x2 = torch.empty(B, 4, h, w, device=dev)
    t_dev = [torch.tensor([float(t)], device=dev) for t in ts]
    coefs = []
    for t in ts:
        a_t, a_p = sch.coefficients(int(t))
        coefs.append(torch.tensor([math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_p), math.sqrt(1 - a_p)], device=dev))
    eng.guidance.fill_(4.5)

    def one_step(i):
        k = i % len(ts)
        torch.cat([lat, lat], out=x2)
        eng.step(x2, t_dev[k], (1.0,) * 6)
        eng.coef.copy_(coefs[k])
        ops.cfg_ddim(eng.eps_out, lat, eng.guidance, eng.coef)

    for i in range(args.warmup):
        if i % len(ts) == 0:
            lat.copy_(host["latents"], non_blocking=True)
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        if i % len(ts) == 0:
            lat.copy_(host["latents"], non_blocking=True)  # a new image starts every 20 steps
        one_step(i)
    if world > 1:  # the only collective on the path: gather final latents (32 KB per row)
        from edgestyle_b200.dist import gather_latents

        gathered = gather_latents(lat, world * images, rank, world)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = tms.item()
    clocks = sampler.stop() if rank == 0 else None
    launches_timed = eng.launches_per_step * args.steps + (ops.LAUNCHES - l0)  # graph replays + eager cfg_ddim
    ms_per_step = ms / args.steps
    value = world * args.steps / (ms / 1e3)

    # ---------------- e2e: the public pipeline call with HOST (pinned) inputs, host result ----------------
    def e2e_call(nsteps):
        out = pipe(image=host["conds"], prompt_embeds=host["prompt_embeds"],
                   negative_prompt_embeds=host["negative_prompt_embeds"], latents=host["latents"],
                   num_inference_steps=nsteps, guidance_scale=4.5, output_type="latent",
                   callback_on_step_end=step_readback)
        res = out.images.to("cpu")
        return res

    d2h = [0]
    host_lat = torch.empty(images, 4, h, w).pin_memory()

    def step_readback(p, i, t, kw):
        host_lat.copy_(kw["latents"], non_blocking=False)  # per-step device->host read of the step's result
        d2h[0] += host_lat.numel() * 4
        return {}

    e2e_steps = 20
    e2e_call(e2e_steps)  # warm-up (graph already captured)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    d2h[0] = 0
    reps = max(1, args.steps // e2e_steps)
    e0.record()
    for _ in range(reps):
        e2e_call(e2e_steps)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = tms.item()
    e2e_value = world * reps * e2e_steps / (e2e_ms / 1e3)
    h2d_per_step = pipe.h2d_bytes / e2e_steps
    d2h_per_step = (d2h[0] / reps + images * 4 * h * w * 4) / e2e_steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_hbm, peak_src = measured_peaks()
    achieved_tf = TFLOP_PER_STEP_B2 * images / (ms_per_step / 1e3) if ms_per_step > 0 else 0.0
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 (fp32 accumulate / statistics)", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: 20-step DDIM sampling 512x512 (64x64 latent), guidance 4.5, "
                               f"CFG batch 2 x {images} image(s) per GPU, SD1.5 UNet + 6 ControlNet/ControlLoRA(rank 32) "
                               "branches + EdgeStyle merge, random-init weights",
                   "rows_per_gpu": B, "latent": [h, w], "parallelism": f"dp{world} (replica per GPU, NCCL gather of latents)",
                   "l2": "per-step weight working set ~3.4 GB >> 126 MB L2, no explicit flush",
                   "cuda_graph": True,
                   "cached_across_steps": "cross-attention K/V projections of the prompt (0.46 % of reference FLOPs, "
                                          "not credited in roofline.achieved)"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_per_step),
                "d2h_bytes_per_step": int(d2h_per_step),
                "what": "EdgeStyleStableDiffusionControlNetPipeline.__call__ (20 steps) from pinned host tensors "
                        "(prompt embeds, 6 cached cond embeddings, latents) to host latents, per-step latent read-back"},
        "gpu_launches": int(launches_timed),
        "roofline": {"bound": "tensor", "achieved": round(achieved_tf, 2), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(achieved_tf / peak_tf, 4),
                     "traffic": DRAM_BYTES_PER_STEP_B2 * images if images == 1 else None,
                     "kernel": "es::gemm_kernel (tcgen05 implicit GEMM) -- whole-step algorithmic FLOPs "
                               f"({TFLOP_PER_STEP_B2} TFLOP per CFG pair, SURVEY.md 8(d)) over the CUDA-event step time; "
                               "traffic = DRAM bytes of all launches of one step (profiles/r1c_launches_final.csv)",
                     "peak_source": peak_src},
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(sample_steps=3)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(sample_steps: int = 1, warmup: int = 0):
    """The fp32 oracle (pure-PyTorch restatement of the reference step) on the host cores."""
    from oracle.schedulers import DDIMScheduler
    from oracle.sd15 import SD15Config
    from oracle.step import build_models, cfg_combine, fused_step, synthetic_inputs

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = SD15Config()
    m = build_models(cfg, (64, 64), rank=32)
    inp = synthetic_inputs(cfg, 1, 64, 64)
    sch = DDIMScheduler()
    ts = sch.set_timesteps(20)
    lat = inp.latents.clone()
    times = []
    with torch.no_grad():
        for i in range(warmup + sample_steps):
            t0 = time.perf_counter()
            x = torch.cat([lat] * 2)
            eps = fused_step(m, x, ts[i % 20], inp.prompt_embeds, inp.conditioning_scale, inp.conds)
            lat = sch.step(cfg_combine(eps, 4.5), ts[i % 20], lat)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": round(1.0 / sec, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_steps} denoise step(s) of the same workload (B=2, 64x64 latent, fp32 PyTorch oracle, "
                      f"{warmup} warm-up), {sec:.2f} s/step"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself cannot be installed
    (needs diffusers==0.26.3, absent offline), so this times the oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    budget_steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    t0 = time.perf_counter()
    cb = cpu_baseline(sample_steps=budget_steps, warmup=warm)
    wall = time.perf_counter() - t0
    v = cb["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": budget_steps,
        "warmup": warm, "ms_per_step": round(1e3 / v, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] step on the host CPU (fp32 oracle port of the reference step); "
                               f"bounded sample: {budget_steps} of the requested {args.steps} steps", "rows_per_gpu": 2,
                   "latent": [64, 64], "parallelism": "host threads"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(wall, 1),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=1, help="images (CFG pairs) per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
