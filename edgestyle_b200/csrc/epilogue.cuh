// Elementwise epilogue arithmetic of the GEMM kernels on packed fp32 pairs (FFMA2 / FADD2 / FMUL2 of sm_100).
// The epilogues are bound by instruction ISSUE -- two warps per SM sub-partition, ~65 % fp32 instructions
// (tools/persist_trace.py) -- and a packed instruction takes one issue slot for two lanes
// (tools/ubench_f32x2.cu: the fp32 rate itself does not change).  tcgen05.ld delivers an accumulator row in consecutive
// registers, so (a[2j], a[2j+1]) are register pairs already.
#pragma once
#include "ptx.cuh"

namespace es {

__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t bc2(float v) { return pk2(v, v); }

// a[0..16) += v[0..16)   (v: 16-byte aligned shared memory)
__device__ __forceinline__ void epi_add_vec16(float (&a)[16], const float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b = *reinterpret_cast<const float4*>(v + 4 * j);
    upk2(add2(pk2(a[4 * j], a[4 * j + 1]), pk2(b.x, b.y)), a[4 * j], a[4 * j + 1]);
    upk2(add2(pk2(a[4 * j + 2], a[4 * j + 3]), pk2(b.z, b.w)), a[4 * j + 2], a[4 * j + 3]);
  }
}
// LayerNorm folded around the GEMM: a = (a - mean * colsum) * rstd + bias
__device__ __forceinline__ void epi_ln_vec16(float (&a)[16], const float* bias, const float* colsum, float mean, float rstd) {
  const uint64_t nm2 = bc2(-mean), r2 = bc2(rstd);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * j);
    const float4 f = *reinterpret_cast<const float4*>(colsum + 4 * j);
    uint64_t t = fma2(pk2(f.x, f.y), nm2, pk2(a[4 * j], a[4 * j + 1]));
    upk2(fma2(t, r2, pk2(b.x, b.y)), a[4 * j], a[4 * j + 1]);
    t = fma2(pk2(f.z, f.w), nm2, pk2(a[4 * j + 2], a[4 * j + 3]));
    upk2(fma2(t, r2, pk2(b.z, b.w)), a[4 * j + 2], a[4 * j + 3]);
  }
}
__device__ __forceinline__ void epi_scale16(float (&a)[16], float alpha) {
  if (alpha == 1.0f) return;  // (uniform)
  const uint64_t s2 = bc2(alpha);
#pragma unroll
  for (int j = 0; j < 8; ++j) upk2(mul2(pk2(a[2 * j], a[2 * j + 1]), s2), a[2 * j], a[2 * j + 1]);
}
// o += residual (8 packed 16-bit pairs)
template <typename T>
__device__ __forceinline__ void epi_add_res16(float (&o)[16], const uint32_t (&rr)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 f = Cvt<T>::unpack2(rr[j]);
    upk2(add2(pk2(o[2 * j], o[2 * j + 1]), pk2(f.x, f.y)), o[2 * j], o[2 * j + 1]);
  }
}
// (sum, sum of squares) of a row, two interleaved partial sums each
__device__ __forceinline__ void epi_rowstat16(const float (&o)[16], uint64_t& rs2, uint64_t& rq2) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint64_t v = pk2(o[2 * j], o[2 * j + 1]);
    rs2 = add2(rs2, v);
    rq2 = fma2(v, v, rq2);
  }
}
// o = alpha * a * gelu(g) over 16 columns; exact (erf) GELU through Abramowitz-Stegun 7.1.26 as in gelu_erf_f (ptx.cuh),
// polynomial coefficients negated so that erf|x| = 1 + (p' t) e: per pair 2 rcp + 2 ex2 on the MUFU pipe, 2 FMUL (|x| is a
// free source modifier there), 2 LOP3 (copysign) and 11 packed instructions.  Written STAGE BY STAGE across the eight pairs: the epilogue has two warps per
// sub-partition and nothing else to hide a dependency chain behind -- pair after pair (rcp -> 4 dependent FMAs -> ex2 ->
// ...) the sixteen GELUs of a chunk took ~1000 cycles (tools/persist_trace.py with the GELU ablated), eight chains wide
// every instruction has seven independent neighbours.
__device__ __forceinline__ void epi_geglu16(float (&o)[16], const float (&a)[16], const float (&g)[16], float alpha) {
  uint64_t ax2[8], t2[8], p2[8], e2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    ax2[j] = pk2(fabsf(g[2 * j]) * 0.70710678118654752f, fabsf(g[2 * j + 1]) * 0.70710678118654752f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float d0, d1;
    upk2(fma2(ax2[j], bc2(0.3275911f), bc2(1.0f)), d0, d1);
    t2[j] = pk2(__fdividef(1.0f, d0), __fdividef(1.0f, d1));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float m0, m1, e0, e1;
    upk2(mul2(mul2(ax2[j], bc2(-1.4426950408889634f)), ax2[j]), m0, m1);  // -ax^2 log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(m0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(m1));
    e2[j] = pk2(e0, e1);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) p2[j] = fma2(t2[j], bc2(-1.061405429f), bc2(1.453152027f));
#pragma unroll
  for (int j = 0; j < 8; ++j) p2[j] = fma2(p2[j], t2[j], bc2(-1.421413741f));
#pragma unroll
  for (int j = 0; j < 8; ++j) p2[j] = fma2(p2[j], t2[j], bc2(0.284496736f));
#pragma unroll
  for (int j = 0; j < 8; ++j) p2[j] = fma2(p2[j], t2[j], bc2(-0.254829592f));
#pragma unroll
  for (int j = 0; j < 8; ++j) p2[j] = fma2(mul2(p2[j], t2[j]), e2[j], bc2(1.0f));  // erf|x|
  const uint64_t h2 = bc2(0.5f * alpha);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float r0, r1;
    upk2(p2[j], r0, r1);
    const uint64_t xh2 = mul2(pk2(g[2 * j], g[2 * j + 1]), h2);                                 // alpha x / 2
    const uint64_t gl = fma2(xh2, pk2(copysignf(r0, g[2 * j]), copysignf(r1, g[2 * j + 1])), xh2);  // alpha gelu(x)
    upk2(mul2(pk2(a[2 * j], a[2 * j + 1]), gl), o[2 * j], o[2 * j + 1]);
  }
}

}  // namespace es
