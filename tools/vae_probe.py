"""CUDA-event timing of the per-call VAE stages at the reference's operating point:
encode of the three ControlLoRA control images (512x512; agnostic + two outfits, edgestyle_pipeline.py:660-662) and
decode of one 64x64 latent (:552-557), real SD1.5 VAE widths, random-init weights.

    python tools/vae_probe.py [--hw 512] [--reps 5]

Prints issued tensor-core FLOPs (2 * M * N * K summed over the es_gemm launches of one call), launches, ms, TFLOP/s."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    from edgestyle_b200 import ops
    from edgestyle_b200 import vae as V

    cfg = V.VaeConfig()
    g = torch.Generator().manual_seed(0)
    sd = {}
    for k, shp in V.vae_spec(cfg).items():
        if "norm" in k:
            sd[k] = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
        elif k.endswith("bias"):
            sd[k] = torch.zeros(shp)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / fan_in ** 0.5
    vae = V.AutoencoderKL(cfg, sd)
    flops = [0]
    real_gemm = ops.gemm

    def counting_gemm(a, b, n, **kw):
        flops[0] += 2 * a.shape[0] * n * b.shape[1]
        return real_gemm(a, b, n, **kw)

    x = (torch.rand(3, 3, args.hw, args.hw, generator=g) * 2 - 1).cuda()
    z = torch.randn(1, 4, args.hw // 8, args.hw // 8, generator=g).cuda()
    for name, fn in (("encode 3 x %d^2 (+ sample)" % args.hw, lambda: vae.encode(x).latent_dist.sample()),
                     ("decode 1 x %d^2" % args.hw, lambda: vae.decode(z).sample)):
        fn()  # warm-up (allocates the scratch, tunes nothing: heuristics only unless a DenoiseEngine enabled the tuner)
        ops.gemm = counting_gemm
        V.ops.gemm = counting_gemm
        flops[0] = 0
        n0 = ops.LAUNCHES if hasattr(ops, "LAUNCHES") else 0
        fn()
        n1 = ops.LAUNCHES if hasattr(ops, "LAUNCHES") else 0
        ops.gemm = real_gemm
        V.ops.gemm = real_gemm
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        print(f"{name}: {flops[0] / 1e12:.3f} TFLOP issued, {n1 - n0} launches, {ms:.2f} ms, "
              f"{flops[0] / ms / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
