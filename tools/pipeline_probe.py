"""Two independent image groups per GPU, each with its own step engine and CUDA graph, replayed on two streams (the
decoder of one group runs under the encoders of the other) against ONE engine over all rows.
    python tools/pipeline_probe.py [images_total]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import config as C  # noqa: E402
from edgestyle_b200.engine import DenoiseEngine  # noqa: E402
from edgestyle_b200.synth import synth_state_dicts  # noqa: E402

images = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = C.UNetConfig()
h = w = 64
sds = synth_state_dicts(cfg, h, w, rank=32, seed=0)
g = torch.Generator().manual_seed(1)


def mk(rows):
    eng = DenoiseEngine(cfg, sds["unet"], sds["lora"], sds["pose"], sds["merge"], rows=rows, h=h, w=w, use_graph=True)
    eng.set_prompt(torch.randn(rows, 77, 768, generator=g))
    eng.set_conditioning([torch.randn(rows, 320, h, w, generator=g) * 0.5 for _ in range(6)])
    x = torch.randn(rows, 4, h, w, generator=g).cuda()
    for _ in range(2):
        eng.step(x, 500.0)
    torch.cuda.synchronize()
    return eng, x


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


one, x1 = mk(2 * images)
t_one = timeit(lambda: one.step(x1, 500.0))
print(f"one engine, {2 * images} rows: {t_one:.3f} ms/step", flush=True)
del one
torch.cuda.empty_cache()
ea, xa = mk(images)
eb, xb = mk(images)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
key = (True,) * 6
ga, gb = ea._graphs[key], eb._graphs[key]


def both():
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(main)
    for st, gr in ((sa, ga), (sb, gb)):
        st.wait_event(ev)
        with torch.cuda.stream(st):
            gr.replay()
        e = torch.cuda.Event()
        e.record(st)
        main.wait_event(e)


t_two = timeit(both)
print(f"two engines x {images} rows on two streams: {t_two:.3f} ms per step of all {2 * images} rows  ({t_one / t_two:.3f} x)", flush=True)
t_a = timeit(lambda: ga.replay())
print(f"(one {images}-row engine alone: {t_a:.3f} ms/step)")
