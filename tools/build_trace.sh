#!/bin/sh
# Builds the instrumented copies of the library used by tools/gemm_trace.py and tools/att_trace.py.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --use_fast_math -shared -cudart shared -DES_PDL_TRIGGER_MAX_CTAS=0 -DES_ATT_POLY_EVERY=0 -DES_PDL_LATE_TRIGGER=0"
S="edgestyle_b200/csrc/api.cu edgestyle_b200/csrc/gemm.cu edgestyle_b200/csrc/attention.cu edgestyle_b200/csrc/norm.cu edgestyle_b200/csrc/merge.cu edgestyle_b200/csrc/elementwise.cu edgestyle_b200/csrc/vae.cu"
nvcc $F -DES_GEMM_TRACE $S -o tools/libgemm_trace.so &
nvcc $F -DES_ATT_TRACE $S -o tools/libatt_trace.so &
wait
