#!/usr/bin/env python
"""bench.py -- EdgeStyle denoise hot path on B200: denoise steps/s at 512x512 (64x64 latent).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--images I]
                  [--config 2|3|4|5] [--sweep 1,2,4,...] [--split-pairs] [--no-lib-baseline] [--no-kernel-rooflines]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--config 2 (default) is the configuration the metric is quoted on; 3 = guidance sweep 3.0/4.5/6.0/7.5 x CFG (4 units
sharded over the ranks, `--split-pairs`: each CFG pair over two GPUs with a per-step noise exchange); 4 = throughput
sweep over rows per GPU (`--sweep`, batch 1 = one row without CFG), one JSON line per batch; 5 = 768x1024 (96x128
latent) in bf16.  Only the default run is the driver's contract line.

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
TFLOP_PER_STEP_B2_96x128 = 19.441 - 0.022  # SURVEY.md 8(d) table, same credit rule
# DRAM bytes of one step (ncu dram__bytes_read.sum + dram__bytes_write.sum over every launch of one CUDA-graph replay,
# caches flushed per kernel, so L2-resident activations count as DRAM reads): read from the committed launch-list
# summary of the CURRENT tree (profiles/step_traffic.json, written by tools/launch_summary.py), never a constant here.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "step_traffic.json")


def step_traffic():
    try:
        d = json.load(open(TRAFFIC_FILE))
        return d.get("dram_bytes_per_step"), d.get("source")
    except Exception:
        return None, None


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


def build_ours(images: int, device, seed: int, h: int = 64, w: int = 64, dtype=torch.float16, n_host_images=None):
    """Synthetic seeded weights (random init of the SD1.5 + EdgeStyle architecture) and pinned host inputs."""
    from edgestyle_b200 import config as C
    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      EdgeStyleStableDiffusionControlNetPipeline, UNet2DConditionModel)
    from edgestyle_b200.synth import synth_state_dicts

    cfg = C.UNetConfig()
    sds = synth_state_dicts(cfg, h, w, rank=32, seed=seed, device="cpu")
    unet = UNet2DConditionModel(cfg, sds["unet"])
    agn = ControlLoRAModel(cfg, sds["lora"][0], lora_linear_rank=32, unet=unet)
    clo = ControlLoRAModel(cfg, sds["lora"][1], lora_linear_rank=32, unet=unet)
    pose = CachedControlNetModel(cfg, sds["pose"])
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], sds["merge"], (h, w), dtype=dtype)
    pipe = EdgeStyleStableDiffusionControlNetPipeline(unet=unet, controlnet=multi, use_graph=True)
    g = torch.Generator().manual_seed(1234 + seed)
    n = n_host_images or images
    host = {
        "latents": torch.randn(n, 4, h, w, generator=g).pin_memory(),
        "prompt_embeds": torch.randn(n, 77, 768, generator=g).pin_memory(),
        "negative_prompt_embeds": torch.randn(n, 77, 768, generator=g).pin_memory(),
    }
    # cached conditioning embeddings: one per image, CFG-duplicated like prepare_image does (edgestyle_pipeline.py:657-658)
    per_img = [(torch.randn(n, 320, h, w, generator=g) * 0.5) for _ in range(6)]
    host["conds_per_image"] = [c.pin_memory() for c in per_img]
    host["conds"] = [torch.cat([c, c]).pin_memory() for c in per_img]
    return cfg, pipe, multi, host, (h, w)


def kernel_rooflines(dev):
    """Per-kernel roofline entries of the step's dominant kernels at step shapes: CUDA-event time of a CUDA graph of `reps`
    back-to-back launches (warm L2, outside the step's timed region), algorithmic FLOPs (2 M N K; attention 4 B h Nq Nkv d) or bytes
    against the measured peaks.  Shapes: BASELINE config 2 base pass (8 images)."""
    from edgestyle_b200 import ops

    peak_tf, peak_hbm, _ = measured_peaks()
    burst = peak_tf
    try:
        burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", peak_tf)
    except Exception:
        pass
    out = []
    reps = 10

    def timeit(fn):
        # `reps` launches captured into one CUDA graph: an eager loop would time the HOST launch rate (~40 us per call
        # through ctypes + tensor-map encoding) for every kernel shorter than that
        fn()
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for _ in range(reps):
                fn()
        gph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    f16 = dict(device=dev, dtype=torch.float16)
    for (imgs, hw, cin, cout) in [(8, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (8, 8, 1280, 1280)]:
        M = imgs * hw * hw
        x = torch.randn(M, cin, **f16)
        wt = torch.randn(cout, 9 * cin, **f16) * (9 * cin) ** -0.5
        o = torch.empty(M, cout, **f16)
        bias = torch.zeros(cout, device=dev)
        us = timeit(lambda: ops.gemm(x, wt, cout, out=o, taps=9, whn=(hw, hw, imgs), bias=bias, c1=cin))
        fl = 2.0 * M * cout * 9 * cin
        out.append({"kernel": "es_gemm conv3x3", "shape": f"{imgs}x{hw}x{hw} {cin}->{cout}", "us": round(us, 1),
                    "bound": "tensor", "achieved": round(fl / us / 1e6, 1), "unit": "TFLOP/s", "frac": round(fl / us / 1e6 / burst, 3)})
    for (M, N, K, geglu) in [(32768, 2560, 320, True), (32768, 960, 320, False), (32768, 320, 1280, False), (8192, 5120, 640, True)]:
        x = torch.randn(M, K, **f16)
        wt = torch.randn(N, K, **f16) * K ** -0.5
        o = torch.empty(M, N // 2 if geglu else N, **f16)
        bias = torch.zeros(N, device=dev)
        # the tile width the engine's tuner would choose among the kernels that can run the layer (GEGLU weights are
        # packed for 160-column sub-tiles: one-tile 160, CTA-pair 320, persistent 1160)
        us, bn_used = None, 0
        for bn in ((160, 320, 1160) if geglu else (0, 320, 1160, 1256)):
            try:
                t = timeit(lambda: ops.gemm(x, wt, N, out=o, bias=bias, act=1 if geglu else 0, block_n=bn))
            except Exception:
                continue
            if us is None or t < us:
                us, bn_used = t, bn
        fl = 2.0 * M * N * K
        out.append({"kernel": "es_gemm linear" + (" + GEGLU" if geglu else ""), "shape": f"M={M} N={N} K={K} (block_n {bn_used})", "us": round(us, 1),
                    "bound": "tensor", "achieved": round(fl / us / 1e6, 1), "unit": "TFLOP/s", "frac": round(fl / us / 1e6 / burst, 3)})
    for (batch, heads, d, n) in [(8, 8, 40, 4096), (8, 8, 80, 1024), (8, 8, 160, 256)]:
        Cc = heads * d
        qkv = torch.randn(batch * n, 3 * Cc, **f16)
        o = torch.empty(batch * n, Cc, **f16)
        us = timeit(lambda: ops.attention(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], o, batch, heads, n, n))
        fl = 4.0 * batch * heads * n * n * d
        # the softmax needs one MUFU exponential per score: 16 / clk / SM -> the binding roofline at these head dims
        ex = batch * heads * n * n
        mufu_us = ex / (148 * 16 * 1.965e9) * 1e6
        out.append({"kernel": "es_attention", "shape": f"b={batch} h={heads} d={d} n={n}", "us": round(us, 1), "bound": "mufu",
                    "achieved": round(fl / us / 1e6, 1), "unit": "TFLOP/s", "frac": round(mufu_us / us, 3),
                    "frac_tensor": round(fl / us / 1e6 / burst, 3)})
    # bandwidth kernels: GroupNorm apply (read + write), merge (>= 295.6 MB per step at 64x64)
    x = torch.randn(32768, 320, **f16)
    o = torch.empty_like(x)
    ws = torch.zeros(8, 32, 2, device=dev)
    ws[..., 1] = 10.0 * 4096  # (sum, sumsq) of unit-variance data: 10 channels per group x 4096 pixels
    gamma, beta = torch.ones(320, device=dev), torch.zeros(320, device=dev)
    us = timeit(lambda: ops.groupnorm(x, o, gamma, beta, ws, 8, 4096, 32, 1e-5, True, stats_ready=True))
    by = 2.0 * x.numel() * 2
    out.append({"kernel": "es_groupnorm_apply", "shape": "32768 x 320", "us": round(us, 1), "bound": "hbm",
                "achieved": round(by / us / 1e3, 1), "unit": "GB/s", "frac": round(by / us / 1e3 / peak_hbm, 3)})
    return out


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
    if args.config == 3:
        return run_config3(args, rank, world, dev)
    h, w = (96, 128) if args.config == 5 else (64, 64)
    dtype = torch.bfloat16 if args.config == 5 else torch.float16
    tflop_pair = TFLOP_PER_STEP_B2_96x128 if args.config == 5 else TFLOP_PER_STEP_B2
    sweep = [int(x) for x in args.sweep.split(",")] if args.sweep else [None]
    for rows in sweep:
        if rows is None:
            images, cfg_on = args.images, True
        else:
            images, cfg_on = max(rows // 2, 1), rows > 1  # batch 1 = one row without CFG (guidance <= 1)
        line = bench_one(args, rank, world, dev, images, cfg_on, h, w, dtype, tflop_pair, first=(rows == sweep[0]))
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_one(args, rank, world, dev, images, cfg_on, h, w, dtype, tflop_pair, first=True):
    import torch.distributed as dist

    from edgestyle_b200 import ops
    from edgestyle_b200.schedulers import DDIMScheduler

    cfg, pipe, multi, host, _ = build_ours(images, dev, seed=rank, h=h, w=w, dtype=dtype)
    B = 2 * images if cfg_on else images
    eng = multi.engine(B, h, w, use_graph=True)
    sch = DDIMScheduler()
    ts = sch.set_timesteps(20)
    # ---------------- device-resident timed region: K steps of (graph replay + CFG/DDIM) ----------------
    pe = torch.cat([host["negative_prompt_embeds"], host["prompt_embeds"]]) if cfg_on else host["prompt_embeds"]
    eng.set_prompt(pe.to(dev))
    eng.set_conditioning([(c if cfg_on else c[:images]).to(dev) for c in host["conds"]])
    lat = host["latents"].to(dev).clone()
    x2 = torch.empty(B, 4, h, w, device=dev)
    t_dev = [torch.tensor([float(t)], device=dev) for t in ts]
    coefs, lin = [], []
    for t in ts:
        a_t, a_p = sch.coefficients(int(t))
        al, sg, alp, sgp = math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_p), math.sqrt(1 - a_p)
        coefs.append(torch.tensor([al, sg, alp, sgp], device=dev))
        lin.append((alp / al, sgp - alp * sg / al))
    eng.guidance.fill_(4.5)

    def one_step(i):
        k = i % len(ts)
        if cfg_on:
            torch.cat([lat, lat], out=x2)
            eng.step(x2, t_dev[k], (1.0,) * 6)
            eng.coef.copy_(coefs[k])
            ops.cfg_ddim(eng.eps_out, lat, eng.guidance, eng.coef)
        else:  # one row per image, no CFG: x' = (a'/a) x + (s' - a' s / a) eps
            eng.step(lat, t_dev[k], (1.0,) * 6)
            ops.lincomb(lat, [(lin[k][0], lat), (lin[k][1], eng.eps_out)])

    for i in range(args.warmup):
        if i % len(ts) == 0:
            lat.copy_(host["latents"], non_blocking=True)
        one_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
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

        gathered = gather_latents(lat, world * images, rank, world)  # noqa: F841
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
    d2h = [0]
    # per-step device->host read of the step's result into two pinned buffers: the copy of step i is enqueued behind
    # the step (stream order: it reads the latents before step i + 1 updates them in place) and the host waits for it one
    # step later, when it reuses the buffer -- a blocking read would stall the launch of the next step behind every step
    host_lat = [torch.empty(images, 4, h, w).pin_memory() for _ in range(2)]
    landed = [None, None]

    def step_readback(p, i, t, kw):
        k = i & 1
        if landed[k] is not None:
            landed[k].synchronize()
        host_lat[k].copy_(kw["latents"], non_blocking=True)
        landed[k] = torch.cuda.Event()
        landed[k].record()
        d2h[0] += host_lat[k].numel() * 4
        return {}

    def e2e_call(nsteps):
        # one cached conditioning embedding per image: the CFG duplication happens on the device, where the reference's
        # prepare_image does it (edgestyle_pipeline.py:657-658)
        out = pipe(image=host["conds_per_image"], prompt_embeds=host["prompt_embeds"],
                   negative_prompt_embeds=host["negative_prompt_embeds"] if cfg_on else None, latents=host["latents"],
                   num_inference_steps=nsteps, guidance_scale=4.5 if cfg_on else 1.0, output_type="latent",
                   callback_on_step_end=step_readback)
        return out.images.to("cpu")

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
        return None
    peak_tf, peak_hbm, peak_src = measured_peaks()
    tflop_step = tflop_pair * images * (1.0 if cfg_on else 0.5)
    achieved_tf = tflop_step / (ms_per_step / 1e3) if ms_per_step > 0 else 0.0
    traffic, traffic_src = step_traffic()
    latent = f"{8 * h}x{8 * w} ({h}x{w} latent)"
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": ("bf16" if dtype == torch.bfloat16 else "f16") + " (fp32 accumulate / statistics)",
        "data": "synthetic",
        "config": {"workload": f"BASELINE configs[{args.config - 1}]: 20-step DDIM sampling {latent}, guidance 4.5, "
                               + (f"CFG batch 2 x {images} image(s)" if cfg_on else f"{images} row(s) without CFG")
                               + " per GPU, SD1.5 UNet + 6 ControlNet/ControlLoRA(rank 32) branches + EdgeStyle merge, "
                               "random-init weights",
                   "rows_per_gpu": B, "latent": [h, w], "parallelism": f"dp{world} (replica per GPU, NCCL gather of latents)",
                   "l2": "per-step weight working set ~3.4 GB >> 126 MB L2, no explicit flush",
                   "cuda_graph": True,
                   "cached_across_steps": "cross-attention K/V projections of the prompt (0.46 % of reference FLOPs, "
                                          "not credited in roofline.achieved)"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_per_step),
                "d2h_bytes_per_step": int(d2h_per_step),
                "what": "EdgeStyleStableDiffusionControlNetPipeline.__call__ (20 steps) from pinned host tensors "
                        "(prompt embeds, 6 cached cond embeddings, latents) to host latents, per-step latent read-back (asynchronous into two pinned "
                        "buffers; the final latents are read synchronously before the clock stops)"},
        "gpu_launches": int(launches_timed),
        "launches_per_step": int(eng.launches_per_step) + 1,
        "roofline": {"bound": "tensor", "achieved": round(achieved_tf, 2), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(achieved_tf / peak_tf, 4),
                     "traffic": traffic * images if (traffic and args.config == 2 and cfg_on) else None,
                     "traffic_source": traffic_src,
                     "kernel": "es::gemm_kernel family (tcgen05 implicit GEMM) -- whole-step algorithmic FLOPs "
                               f"({tflop_step:.3f} TFLOP per step, SURVEY.md 8(d)) over the CUDA-event step time",
                     "peak_source": peak_src},
    }
    if first and not args.no_kernel_rooflines and world == 1:
        try:
            line["roofline"]["kernels"] = kernel_rooflines(dev)
        except Exception as exc:  # a probe must never cost the contract line
            line["roofline"]["kernels_error"] = f"{type(exc).__name__}: {exc}"[:200]
    if first and not args.no_lib_baseline and world == 1 and args.config == 2 and cfg_on and images == 1:
        line["gpu_library_baseline"] = gpu_library_baseline_subprocess(h, w, images, dtype)
    if first and not args.no_cpu_baseline and world == 1 and args.config == 2:
        line["cpu_baseline"] = cpu_baseline(sample_steps=3)
    del eng, pipe, multi
    torch.cuda.empty_cache()
    return line


def gpu_library_baseline_subprocess(h, w, images, dtype):
    """tools/lib_baseline.py in its own process (it holds a second, fp16-cast copy of the oracle on the GPU): the stock
    PyTorch dispatch (cuDNN / cuBLAS / SDPA) of the same step, eager and CUDA-graphed -- the bar SURVEY.md 2.1 sets."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "lib_baseline.py"), "--images", str(images), "--hw", str(h), str(w),
           "--dtype", "bf16" if dtype == torch.bfloat16 else "fp16"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)["gpu_library_baseline"]
        return {"error": (out.stderr or "no output")[-300:]}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


def run_config3(args, rank, world, dev):
    """BASELINE configs[2]: guidance-scale sweep 3.0 / 4.5 / 6.0 / 7.5 x CFG = 8 rows = 4 (image, scale) units, sharded
    over the ranks (a unit keeps both CFG rows on one GPU; `--split-pairs`: one CFG pair over two GPUs, per-step noise
    exchange).  value = denoise steps/s of the whole sweep (units x steps / time)."""
    import torch.distributed as dist

    from edgestyle_b200.dist import denoise_split_pairs, denoise_units, guidance_sweep_units

    units = guidance_sweep_units(1, [3.0, 4.5, 6.0, 7.5])
    cfg, pipe, multi, host, (h, w) = build_ours(1, dev, seed=0)  # every rank: the same image, weights and noise
    host = {k: ([c[:1] for c in v] if isinstance(v, list) else v) for k, v in host.items()}
    host["conds"] = host["conds_per_image"]
    nsteps = 20
    split = args.split_pairs

    def run():
        if split:
            return denoise_split_pairs(multi, units, host, nsteps, rank, world)
        return denoise_units(pipe, units, host, nsteps, rank, world)

    reps_w, reps = max(1, args.warmup // nsteps), max(1, args.steps // nsteps)
    for _ in range(reps_w):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lat = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = tms.item()
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        unit_steps = len(units) * nsteps * reps
        peak_tf, _, peak_src = measured_peaks()
        achieved = TFLOP_PER_STEP_B2 * unit_steps / (ms / 1e3)
        line = {"metric": METRIC, "value": round(unit_steps / (ms / 1e3), 3), "unit": "unit-steps/s (one CFG pair x one step)",
                "n_gpus": world, "steps": nsteps * reps, "warmup": nsteps * reps_w, "ms_per_step": round(ms / (nsteps * reps), 4),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 (fp32 accumulate / statistics)",
                "data": "synthetic",
                "config": {"workload": "BASELINE configs[2]: guidance sweep 3.0/4.5/6.0/7.5 x CFG = 8 rows (4 units), 20 DDIM "
                                       "steps, 512x512, through the pipeline call with host tensors",
                           "units": len(units), "parallelism": (f"{world} GPUs, CFG pair split over two GPUs + per-step eps "
                                                                "all-gather (64 KB)") if split else
                           f"units sharded over {world} GPU(s), one NCCL gather of the final latents"},
                "clocks": clocks, "latent_checksum": round(float(lat.double().abs().mean().item()), 6),
                "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": round(achieved / peak_tf / world, 4), "peak_source": peak_src}}
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
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs[] entry (1-based)")
    ap.add_argument("--sweep", default="", help="config 4: comma-separated rows per GPU (1 = one row without CFG)")
    ap.add_argument("--split-pairs", action="store_true", help="config 3: split every CFG pair over two GPUs")
    ap.add_argument("--no-lib-baseline", action="store_true",
                    help="skip the stock PyTorch (cuDNN / cuBLAS / SDPA) timing of the same step (N = 1, config 2 only)")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.config == 4 and not args.sweep:
        args.sweep = "1,2,4,8,16,32,64"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
