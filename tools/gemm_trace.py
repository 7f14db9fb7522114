"""clock64 trace of one GEMM CTA: where does a short-K tile spend its time?  (tools/libgemm_trace.so, -DES_GEMM_TRACE)"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ext  # noqa: E402
ext.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgemm_trace.so")
from edgestyle_b200 import ops  # noqa: E402
lib = ext.load()
lib.es_gemm_trace.restype = C.c_int
lib.es_gemm_trace.argtypes = [C.c_void_p]
names = ["start", "setup done", "first stage full (MMA)", "last MMA issued", "accum visible (epi)", "epilogue done", "dealloc/all units", "kb8 full", "kb24 full",
         "epi: vec staged", "epi: TMEM drained", "epi: panels complete", "epi: stats done"]
for (M, N, K, res, bn, conv) in [(512, 1280, 1280, False, 128, (8, 8, 8)), (512, 1280, 1280, False, 64, (8, 8, 8)), (32768, 320, 320, True, 160, None), (32768, 320, 320, False, 160, None), (32768, 960, 320, False, 256, None),
                                 (32768, 320, 1280, True, 160, None), (2048, 1280, 1280, True, 128, None),
                                 (32768, 320, 320, False, 160, (64, 64, 8)), (8192, 640, 640, False, 160, (32, 32, 8))]:
    taps = 9 if conv else 1
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    b = torch.randn(N, K * taps, device="cuda", dtype=torch.float16)
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    r = torch.randn(M, N, device="cuda", dtype=torch.float16) if res else None
    bias = torch.zeros(N, device="cuda")
    kw = dict(out=out, bias=bias, residual=r, block_n=bn, split_k=(6 if bn == 128 else 3) if M == 512 else 1)
    if conv:
        kw.update(taps=9, whn=conv, c1=K)
    for _ in range(3):
        ops.gemm(a, b, N, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm(a, b, N, **kw)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 64)()
    lib.es_gemm_trace(buf)
    t = list(buf)[:13]
    print(f"M={M} N={N} K={K}x{taps} bn={bn} residual={res}: {e0.elapsed_time(e1) * 50:.1f} us/launch; CTA(1,0) cycles:",
          ", ".join(f"{n}=+{t[i] - t[0]}" for i, n in enumerate(names)))
