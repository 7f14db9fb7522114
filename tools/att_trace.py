"""Pipeline trace of one attention CTA (clock64 stamps), built with -DES_ATT_TRACE into tools/libatt_trace.so."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgestyle_b200 import ext  # noqa: E402

ext.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libatt_trace.so")
lib = C.CDLL(ext.LIB_PATH)
lib.es_attention.restype = C.c_int
lib.es_attention.argtypes = [C.POINTER(ext.EsAttention), C.c_void_p]
batch, heads, d, n = 8, 8, 40, 4096
Cc = heads * d
qkv = torch.randn(batch * n, 3 * Cc, device="cuda", dtype=torch.float16)
out = torch.empty(batch * n, Cc, device="cuda", dtype=torch.float16)
a = ext.EsAttention()
a.dtype = 0
a.q, a.k, a.v, a.out = qkv.data_ptr(), qkv[:, Cc:].data_ptr(), qkv[:, 2 * Cc:].data_ptr(), out.data_ptr()
a.ldq = a.ldk = a.ldv = 3 * Cc
a.ldo = Cc
a.bsq = a.bsk = a.bsv = n * 3 * Cc
a.bso = n * Cc
a.batch, a.heads, a.d, a.nq, a.nkv, a.scale = batch, heads, d, n, n, d ** -0.5
for _ in range(3):
    assert lib.es_attention(C.byref(a), None) == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 256)()
assert lib.es_attention_trace(buf) == 0
t = list(buf)
mma, t8, t9 = t[:64], t[64:128], t[128:192]
t0 = min(x for x in mma if x > 0)
# staggered issue order (attention2.cuh, QT == 2): per key tile j the MMA thread stamps after [A] QK0(j+1), [B] PV1(j-1)
# (j == 0: QK1(0)), [C] QK1(j+1), [D] PV0(j)
print("tile | MMA thread: A (QK0(j+1) issued)  B (PV1(j-1) issued)  C (QK1(j+1) issued)  D (PV0(j) issued)   (cycles)")
for j in range(6, 12):
    print(j, [x - t0 for x in mma[4 * j:4 * j + 4]])
for name, arr in (("tile 8", t8), ("tile 9", t9)):
    print(name, "per softmax warp (WG0: warps 2-5, WG1: warps 6-9): s_full s_free exps_done p_full_arrive")
    for w in range(8):
        print("  warp", w + 2, "SMSP", (w + 2) % 4, [v - t0 for v in arr[4 * w:4 * w + 4]])
