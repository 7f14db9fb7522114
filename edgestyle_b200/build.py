"""Build the C-ABI shared library in-tree with nvcc for sm_100a (no torch involvement).

    python -m edgestyle_b200.build [--force]

Output: edgestyle_b200/libedgestyle_b200.so (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libedgestyle_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "gemm.cu", "attention.cu", "norm.cu", "merge.cu", "elementwise.cu", "vae.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# --use_fast_math (approximate division / sqrt, flush-to-zero) only where it was measured to matter: the GEMM epilogues
# and the attention softmax.  The scheduler updates (cfg_ddim / cfg_x0 divide by alpha; latents are carried over 20-50
# steps), the GroupNorm / merge-LayerNorm statistics and the VAE softmax / Gaussian sampling compile with IEEE
# arithmetic; they call __expf explicitly where a fast exponential is wanted (silu_f).
FAST_MATH_SOURCES = {"gemm.cu", "attention.cu"}
# every kernel whose grid has at most this many CTAs lets its stream successor start its prologue (barrier init, TMEM
# allocation, weight-tile loads) while it is still running; 0 disables the early trigger
PDL_TRIGGER_MAX_CTAS = int(os.environ.get("ES_PDL_TRIGGER_MAX_CTAS", "0"))
NVCC_FLAGS.append(f"-DES_PDL_TRIGGER_MAX_CTAS={PDL_TRIGGER_MAX_CTAS}")
# fraction of the softmax exponentials taken off the MUFU pipe (attention.cu): 0 = none, n = every n-th score
NVCC_FLAGS.append(f"-DES_ATT_POLY_EVERY={int(os.environ.get('ES_ATT_POLY_EVERY', '0'))}")
if os.environ.get("ES_ATT2_POLY_MASK"):  # which score pairs of the softmax take the FMA-pipe exponential (attention2.cuh)
    NVCC_FLAGS.append(f"-DES_ATT2_POLY_MASK={int(os.environ['ES_ATT2_POLY_MASK'], 0)}")
NVCC_FLAGS.append(f"-DES_PDL_LATE_TRIGGER={int(os.environ.get('ES_PDL_LATE_TRIGGER', '0'))}")
# resident CTAs per SM the merge's phase 2 is compiled for (1: 222 registers, no spills; 2: 128 registers, 196 B spilled)
NVCC_FLAGS.append(f"-DES_MERGE_P2_BLOCKS={int(os.environ.get('ES_MERGE_P2_BLOCKS', '2'))}")
NVCC_FLAGS += os.environ.get("ES_EXTRA_NVCC_FLAGS", "").split()  # experiments: extra -D switches for a side build
OUT = os.environ.get("ES_LIB_OUT", OUT)
OBJ = os.environ.get("ES_OBJ_DIR", OBJ)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(src_paths, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_paths)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "edgestyle_b200.h"))
    nvcc = _nvcc()

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer([s] + headers, o):
            cmd = [nvcc] + NVCC_FLAGS + (["--use_fast_math"] if src in FAST_MATH_SOURCES else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
            return o, True
        return o, False

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(OUT):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
