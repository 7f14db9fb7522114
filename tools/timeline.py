"""Op-by-op timeline of the CAPTURED multi-stream denoise step (the thing bench.py times).

A 1-thread probe kernel (es_stamp: griddepcontrol.wait, then %globaltimer) is enqueued after every native op on the
stream the op ran on, and captured into the CUDA graph with the step.  After a replay the stamps give, per stream,
the completion time of every op under the real concurrency of the step.

    python tools/timeline.py [images] > gpurun_out/timeline.txt
"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import config as C, ops  # noqa: E402
from edgestyle_b200.engine import DenoiseEngine  # noqa: E402
from edgestyle_b200.ext import load  # noqa: E402
from edgestyle_b200.synth import synth_state_dicts  # noqa: E402

images = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = C.UNetConfig()
h = w = 64
sds = synth_state_dicts(cfg, h, w, rank=32, seed=0)
eng = DenoiseEngine(cfg, sds["unet"], sds["lora"], sds["pose"], sds["merge"], rows=2 * images, h=h, w=w, use_graph=True)
g = torch.Generator().manual_seed(1)
eng.set_prompt(torch.randn(2 * images, 77, 768, generator=g))
eng.set_conditioning([torch.randn(2 * images, 320, h, w, generator=g) * 0.5 for _ in range(6)])
x = torch.randn(2 * images, 4, h, w, generator=g).cuda()

stamps = torch.zeros(4096, dtype=torch.int64, device="cuda")
log = []  # (slot, stream, name, desc, flops)
lib = load()


def stamp(name, desc, flops=0.0):
    slot = len(log)
    st = torch.cuda.current_stream().cuda_stream
    lib.es_stamp(stamps.data_ptr() + 8 * slot, st)
    log.append((slot, st, name, desc, flops))


def wrap(name, fn, describe):
    def inner(*a, **k):
        r = fn(*a, **k)
        if torch.cuda.is_current_stream_capturing():
            desc, fl = describe(*a, **k)
            stamp(name, desc, fl)
        return r
    return inner


def d_gemm(a, b, n, **k):
    M = a.shape[0]
    taps = k.get("taps", 1)
    c1 = k.get("c1") or a.shape[1]
    K = taps * c1 + (k["a2"].shape[1] if k.get("a2") is not None else 0)
    kind = "conv3x3" if taps == 9 else "gemm"
    extra = ("+lora" if k.get("a2") is not None and taps == 1 else "") + ("+sc" if k.get("a2") is not None and taps == 9 else "") \
        + ("+geglu" if k.get("act") else "") + ("+res" if k.get("residual") is not None else "") + ("+gn" if k.get("gn_ws") is not None else "")
    return f"{kind}{extra} M={M} N={n} K={K}", 2.0 * M * n * K


def d_att(q, kk, v, out, batch, heads, nq, nkv, scale=None):
    d = q.shape[1] // heads
    return f"attn b={batch} nq={nq} nkv={nkv} d={d}", 4.0 * batch * heads * nq * nkv * d


ops.gemm = wrap("gemm", ops.gemm, d_gemm)
ops.attention = wrap("attention", ops.attention, d_att)
ops.groupnorm = wrap("groupnorm", ops.groupnorm, lambda x0, out, *a, **k: (f"M={x0.shape[0]} C={out.shape[1]} {'apply' if k.get('stats_ready') else 'stats+apply'}", 0.0))
ops.layernorm = wrap("layernorm", ops.layernorm, lambda x, out, *a, **k: (f"M={x.shape[0]} C={x.shape[1]}", 0.0))
ops.merge_levels = wrap("merge", ops.merge_levels, lambda levels, scale_dev, B: (f"B={B} levels={len(levels)} (3 launches)", 0.0))
for nm in ("small_linear", "im2col3x3", "upsample2x", "nchw_to_nhwc", "timestep_embedding"):
    setattr(ops, nm, wrap(nm, getattr(ops, nm), lambda *a, **k: ("", 0.0)))

# phase markers from the engine's structure: wrap _encoder and record a stamp at its end on its stream
_enc = eng._encoder


def enc_wrapped(E, x_, imgs, temb, ctx, seg, tag, on_level=None):
    if torch.cuda.is_current_stream_capturing():
        stamp("PHASE", f"begin encoder {tag}")
    capturing = torch.cuda.is_current_stream_capturing()
    n_levels = len(eng.res_shapes)

    def on_level2(li, t):
        # the "end" probe must precede the event that joins this stream back into the graph
        if capturing and li == n_levels - 1:
            stamp("PHASE", f"end encoder {tag}")
        if on_level is not None:
            on_level(li, t)

    r = _enc(E, x_, imgs, temb, ctx, seg, tag, on_level2)
    return r


eng._encoder = enc_wrapped

for _ in range(3):
    eng.step(x, 500.0)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(x, 500.0)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
st = stamps[: len(log)].cpu().tolist()
t0 = min(st)
print(f"# images {images}; step with probes {sorted(ts)[2]:.3f} ms; {len(log)} probes; span {(max(st) - t0) / 1e6:.3f} ms")
streams = []
for _, s, *_ in log:
    if s not in streams:
        streams.append(s)
last = {}
fam = collections.defaultdict(lambda: collections.defaultdict(float))
rows = []
for (slot, s, name, desc, fl), t in zip(log, st):
    si = streams.index(s)
    prev = last.get(si)
    dt = (t - prev) / 1e3 if prev is not None else 0.0
    last[si] = t
    rows.append((si, (t - t0) / 1e3, dt, name, desc, fl))
    fam[si][name] += dt
for si in range(len(streams)):
    print(f"# stream {si}: " + ", ".join(f"{k} {v / 1e3:.2f} ms" for k, v in sorted(fam[si].items(), key=lambda kv: -kv[1])))
print("# stream  t_end(us)  dt(us)  TFLOP/s  op")
for si, te, dt, name, desc, fl in rows:
    tf = f"{fl / dt / 1e6:7.1f}" if fl and dt > 0 else "       "
    print(f"{si} {te:10.1f} {dt:8.1f} {tf}  {name} {desc}")
