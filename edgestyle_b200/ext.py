"""ctypes binding of the C-ABI library (include/edgestyle_b200.h).  No fallback: if the library is
missing or a call fails, an exception is raised -- there is deliberately no CPU / PyTorch path here."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ES_LIB", os.path.join(HERE, "libedgestyle_b200.so"))

ES_MAX_SEG = 4
ABI_VERSION = 2  # 2: es_merge_levels (level table, device-side scales) replaced es_merge_phase
DTYPE_F16, DTYPE_BF16 = 0, 1
ACT_NONE, ACT_GEGLU, ACT_SILU = 0, 1, 2

vp = C.c_void_p
ll = C.c_longlong
fp = C.POINTER(C.c_float)


class EsGemm(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("a", vp), ("c1", C.c_int), ("lda", ll), ("w", C.c_int), ("h", C.c_int),
        ("n_img", C.c_int), ("taps", C.c_int), ("b", vp), ("n_total_b", C.c_int), ("a2", vp), ("c2", C.c_int),
        ("lda2", ll), ("b2", vp), ("n_total_b2", C.c_int), ("n", C.c_int), ("nseg", C.c_int),
        ("seg_row_start", C.c_int * (ES_MAX_SEG + 1)), ("seg_b_noff", C.c_int * ES_MAX_SEG),
        ("seg_b2_noff", C.c_int * ES_MAX_SEG), ("bias", vp), ("rowvec", vp), ("rows_per_img", C.c_int),
        ("rowvec_ld", C.c_int), ("residual", vp), ("ldr", ll), ("act", C.c_int), ("alpha", C.c_float),
        ("out", vp), ("ldc", ll), ("out_fp32", C.c_int), ("block_n", C.c_int), ("stages", C.c_int),
        ("split_k", C.c_int), ("b_blocked", C.c_int), ("gn_ws", vp), ("gn_groups", C.c_int), ("gn_cpg", C.c_int), ("gn_col0", C.c_int), ("rowstat_out", vp), ("ln_rowstat", vp), ("ln_colsum", vp),
        ("ln_features", C.c_int), ("ln_eps", C.c_float), ("prefetch", vp), ("prefetch_bytes", ll), ("workspace", vp),
        ("workspace_bytes", ll),
    ]


class EsAttention(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("q", vp), ("k", vp), ("v", vp), ("out", vp), ("ldq", ll), ("ldk", ll), ("ldv", ll),
        ("ldo", ll), ("bsq", ll), ("bsk", ll), ("bsv", ll), ("bso", ll), ("batch", C.c_int), ("heads", C.c_int),
        ("d", C.c_int), ("nq", C.c_int), ("nkv", C.c_int), ("scale", C.c_float),
    ]


class EsGroupNorm(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("x0", vp), ("c0", C.c_int), ("ld0", ll), ("x1", vp), ("c1", C.c_int), ("ld1", ll),
        ("n_img", C.c_int), ("hw", C.c_int), ("groups", C.c_int), ("eps", C.c_float), ("gamma", vp), ("beta", vp),
        ("ws", vp), ("out", vp), ("ldo", ll), ("silu", C.c_int),
    ]


ES_MERGE_MAX_LEVELS = 16


class EsMergeLevel(C.Structure):
    _fields_ = [
        ("res", vp * 6), ("w1", vp), ("b1", vp), ("w2", vp), ("b2", vp), ("w3", vp), ("b3", vp), ("g1", vp),
        ("be1", vp), ("g2", vp), ("be2", vp), ("stats", vp), ("z", vp), ("skip", vp), ("dst", vp), ("gn_ws", vp),
        ("lds", ll), ("ldd", ll), ("hw", C.c_int), ("C", C.c_int), ("gn_groups", C.c_int), ("gn_cpg", C.c_int),
        ("gn_col0", C.c_int), ("z_f32", C.c_int), ("gain", C.c_float),
    ]


class EsMergeBatch(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("B", C.c_int), ("n_levels", C.c_int), ("scale", vp),
        ("levels", EsMergeLevel * ES_MERGE_MAX_LEVELS),
    ]


EXPORTS = {
    "es_last_error": (C.c_char_p, []),
    "es_abi_version": (C.c_int, []),
    "es_set_pdl": (C.c_int, [C.c_int]),
    "es_gemm": (C.c_int, [C.POINTER(EsGemm), vp]),
    "es_attention": (C.c_int, [C.POINTER(EsAttention), vp]),
    "es_groupnorm_stats": (C.c_int, [C.POINTER(EsGroupNorm), vp]),
    "es_groupnorm_apply": (C.c_int, [C.POINTER(EsGroupNorm), vp]),
    "es_groupnorm_fused": (C.c_int, [C.POINTER(EsGroupNorm), vp]),
    "es_layernorm": (C.c_int, [C.c_int, vp, ll, vp, ll, vp, vp, C.c_int, C.c_int, C.c_float, vp]),
    "es_merge_levels": (C.c_int, [C.POINTER(EsMergeBatch), C.c_int, vp]),
    "es_timestep_embedding": (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
    "es_small_linear": (C.c_int, [C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, vp]),
    "es_nchw_to_nhwc": (C.c_int, [C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, ll, vp]),
    "es_nhwc_to_nchw": (C.c_int, [C.c_int, vp, ll, vp, C.c_int, C.c_int, C.c_int, vp]),
    "es_im2col3x3": (C.c_int, [C.c_int, vp, ll, vp, ll, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "es_im2col3x3_pad": (C.c_int, [C.c_int, vp, ll, vp, ll, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, vp]),
    "es_softmax_rows": (C.c_int, [C.c_int, vp, ll, vp, ll, C.c_int, C.c_int, C.c_float, vp]),
    "es_gaussian_sample": (C.c_int, [vp, ll, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp]),
    "es_upsample2x": (C.c_int, [C.c_int, vp, ll, vp, ll, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "es_add": (C.c_int, [C.c_int, vp, ll, vp, ll, vp, ll, C.c_int, C.c_int, vp]),
    "es_cfg_ddim": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]),
    "es_cfg_x0": (C.c_int, [vp, vp, vp, C.c_float, C.c_float, vp, C.c_int, C.c_int, vp]),
    "es_stamp": (C.c_int, [vp, vp]),
    "es_lincomb4": (C.c_int, [vp, C.c_float, vp, C.c_float, vp, C.c_float, vp, C.c_float, vp, ll, vp]),
}

_lib: Optional[C.CDLL] = None


class EdgeStyleNativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the library and type every export.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EdgeStyleNativeError(
            f"{LIB_PATH} not found: build it with `python -m edgestyle_b200.build` (there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.es_abi_version() != ABI_VERSION:
        raise EdgeStyleNativeError("ABI version mismatch between ext.py and libedgestyle_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise EdgeStyleNativeError(f"{what} failed ({rc}): {load().es_last_error().decode()}")
