"""clock64 trace of CTA 0 of the persistent GEMM over its first four units (tools/libgemm_trace.so, -DES_GEMM_TRACE)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ext  # noqa: E402
ext.LIB_PATH = os.environ.get("ES_TRACE_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgemm_trace.so")
from edgestyle_b200 import ops  # noqa: E402
lib = ext.load()
lib.es_gemm_trace.restype = C.c_int
lib.es_gemm_trace.argtypes = [C.c_void_p]
names = ["mma: acc free", "mma: first stage full", "mma: last MMA issued", "epi: (A) panels free", "epi: vec staged",
         "epi: acc visible", "epi: drained", "epi: (C) panels complete", "epi: stats done", "epi: store read done"]
for (M, N, K, res, bn, conv) in [(32768, 2560, 320, "geglu", 160, None), (32768, 320, 320, True, 160, None), (32768, 960, 320, False, 256, None),
                                 (32768, 320, 1280, True, 160, None), (32768, 320, 320, False, 160, (64, 64, 8))]:
    taps = 9 if conv else 1
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    b = torch.randn(N, K * taps, device="cuda", dtype=torch.float16)
    geglu = res == "geglu"
    out = torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=torch.float16)
    r = torch.randn(M, N, device="cuda", dtype=torch.float16) if (res and not geglu) else None
    kw = dict(out=out, bias=torch.zeros(N, device="cuda"), residual=r, block_n=1000 + bn, act=1 if geglu else 0)
    if conv:
        kw.update(taps=9, whn=conv, c1=K)
    for _ in range(3):
        ops.gemm(a, b, N, **kw)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 64)()
    lib.es_gemm_trace(buf)
    t = list(buf)
    t0 = t[16]
    print(f"M={M} N={N} K={K}x{taps} bn={bn} residual={res}")
    for ui in range(4):
        print(f"  unit {ui}: " + ", ".join(f"{n}={t[16 + 12 * ui + k] - t0}" for k, n in enumerate(names)))
