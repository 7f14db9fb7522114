// EdgeStyle ControlNetBlock merge as three bandwidth passes with two per-sample global
// reductions -- for ALL residual levels of the step in ONE launch per pass (a level table in the kernel
// parameters; grid = sum of the levels' CTAs), instead of 13 levels x 3 launches.  The reference's
// stack/permute/contiguous interleave (/root/reference/model/edgestyle_multicontrolnet.py:160-164,479-514)
// is pure indexing here: group g = c*3+p of first_conv pairs nets (2p, 2p+1) of channel c, so the six
// residual slabs are read in place.  conditioning_scale (controllora.py:267-270) is a DEVICE vector, so a
// captured CUDA graph serves every scale; a net whose scale is 0 is not read at all.
#include "common.h"
#include "ptx.cuh"

namespace es {

constexpr int kMergeThreads = 256;
constexpr int kMergeIters = 8;  // pixel rows per thread and CTA
#ifndef ES_MERGE_P2_BLOCKS
#define ES_MERGE_P2_BLOCKS 2
#endif
#define ES_MERGE_P2_BLOCKS_ONLY(PH) ((PH) == 2 ? ES_MERGE_P2_BLOCKS : 2)

template <typename T>
__device__ __forceinline__ void mload8(const T* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = Cvt<T>::unpack2(u.x), b = Cvt<T>::unpack2(u.y), c = Cvt<T>::unpack2(u.z), d = Cvt<T>::unpack2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// streaming variant for data that is read exactly once by this pass
template <typename T>
__device__ __forceinline__ void mload8_cs(const T* p, float (&f)[8]) {
  const uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
  float2 a = Cvt<T>::unpack2(u.x), b = Cvt<T>::unpack2(u.y), c = Cvt<T>::unpack2(u.z), d = Cvt<T>::unpack2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void fload8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

struct MergeLevelK {
  const void* res[6];
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  const void *g1, *be1, *g2, *be2;
  double* stats;
  void* z;
  const void* skip;
  void* dst;
  float* gn_ws;
  long long lds, ldd;
  int hw, C;
  int gn_groups, gn_cpg, gn_col0;
  int z_f32;
  float gain;
  int cta0, chunks, rows_per_cta;
};

struct MergeTable {
  int n_levels, B;
  const float* scale;  // device [6]
  MergeLevelK lv[ES_MERGE_MAX_LEVELS];
};

__device__ __forceinline__ void block_reduce2_to_global(float a, float b, double* dst) {
  __shared__ float red[2][kMergeThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int tid = threadIdx.x;
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = a;
    red[1][tid >> 5] = b;
  }
  __syncthreads();
  if (tid < 32) {
    a = tid < kMergeThreads / 32 ? red[0][tid] : 0.f;
    b = tid < kMergeThreads / 32 ? red[1][tid] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (tid == 0) {
      atomicAdd(dst, static_cast<double>(a));
      atomicAdd(dst + 1, static_cast<double>(b));
    }
  }
}

// grid: sum over levels of B * chunks CTAs; block 256 = (C/8 channel vectors) x ny pixel rows.
// PHASE 1: stats of u.  PHASE 2: z + stats of z.  PHASE 3: output (+ optional GroupNorm statistics of it).
template <typename T, int PHASE>
__global__ void __launch_bounds__(kMergeThreads, ES_MERGE_P2_BLOCKS_ONLY(PHASE)) merge_levels_kernel(const __grid_constant__ MergeTable tb) {
  pdl_launch_dependents();
  int li = 0;
#pragma unroll 1
  for (int i = 1; i < tb.n_levels; ++i)
    if (static_cast<int>(blockIdx.x) >= tb.lv[i].cta0) li = i;
  const MergeLevelK& L = tb.lv[li];
  const int local = blockIdx.x - L.cta0;
  const int b = local / L.chunks;
  const int chunk = local - b * L.chunks;
  const int C = L.C, hw = L.hw;
  const int vpp = C >> 3;
  const int ny = kMergeThreads / vpp;
  const int ty = threadIdx.x / vpp;
  const int v = threadIdx.x - ty * vpp;
  const bool active = ty < ny;
  const int ch = v * 8;
  const int p0 = chunk * L.rows_per_cta;
  const int p1 = min(hw, p0 + L.rows_per_cta);
  const long long img_off = static_cast<long long>(b) * hw * C;
  pdl_wait();  // PDL: inputs are produced by the preceding kernels

  float acc_s = 0.f, acc_q = 0.f;

  if (PHASE == 1 || PHASE == 2) {
    // per-channel first_conv weights ([3 pairs][2 nets][C]) with conditioning_scale (x level gain) folded in
    float wa[3][8], wb[3][8], bb[3][8];
    bool on_a[3], on_b[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float sa = __ldg(tb.scale + 2 * p) * L.gain, sb = __ldg(tb.scale + 2 * p + 1) * L.gain;
      on_a[p] = sa != 0.f;
      on_b[p] = sb != 0.f;
      if (active) {
        fload8(L.w1 + (p * 2 + 0) * C + ch, wa[p]);
        fload8(L.w1 + (p * 2 + 1) * C + ch, wb[p]);
        fload8(L.b1 + p * C + ch, bb[p]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wa[p][j] *= sa;
        wb[p][j] *= sb;
      }
    }
    float mu1 = 0.f, r1 = 0.f;
    float w2[3][8], b2[8];
    if (PHASE == 2) {
      const double n = 3.0 * C * hw;
      const double m = L.stats[b * 4 + 0] / n;
      const double var = L.stats[b * 4 + 1] / n - m * m;
      mu1 = static_cast<float>(m);
      r1 = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
      if (active) {
#pragma unroll
        for (int p = 0; p < 3; ++p) fload8(L.w2 + p * C + ch, w2[p]);
        fload8(L.b2 + ch, b2);
      }
    }
    if (active) {
      // All loads of a pixel (six residual vectors, and in phase 2 the three g1 / be1 pairs) are issued back to back
      // and UNCONDITIONALLY before the first use: with the loads inside `if (scale != 0)` blocks the compiler kept
      // them in six dependent groups and the pass ran at 12 % of the DRAM peak (ncu, profiles/r2p).  A gated net's slab
      // points at valid memory (the engine passes the UNet rows); its values are dropped with a select, not by 0 * x.
      const T* rp[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) rp[k] = reinterpret_cast<const T*>(L.res[k]);
#pragma unroll 2
      for (int p = p0 + ty; p < p1; p += ny) {
        const long long off = img_off + static_cast<long long>(p) * C + ch;
        uint4 rv[6], gv[3], bv[3];
#pragma unroll
        for (int k = 0; k < 6; ++k) rv[k] = __ldg(reinterpret_cast<const uint4*>(rp[k] + off));
        if (PHASE == 2) {
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const long long goff = (static_cast<long long>(p) * 3 + pr) * C + ch;
            gv[pr] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(L.g1) + goff));
            bv[pr] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(L.be1) + goff));
          }
        }
        auto unpack8 = [](const uint4& u, float (&f)[8]) {
          const float2 a = Cvt<T>::unpack2(u.x), b2 = Cvt<T>::unpack2(u.y), c2 = Cvt<T>::unpack2(u.z), d2 = Cvt<T>::unpack2(u.w);
          f[0] = a.x; f[1] = a.y; f[2] = b2.x; f[3] = b2.y; f[4] = c2.x; f[5] = c2.y; f[6] = d2.x; f[7] = d2.y;
        };
        float zacc[8];
        if (PHASE == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) zacc[j] = b2[j];
        }
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          float ra[8], rb[8];
          unpack8(rv[2 * pr], ra);
          unpack8(rv[2 * pr + 1], rb);
          float u[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            u[j] = fmaf(wa[pr][j], on_a[pr] ? ra[j] : 0.f, fmaf(wb[pr][j], on_b[pr] ? rb[j] : 0.f, bb[pr][j]));
          if (PHASE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc_s += u[j];
              acc_q = fmaf(u[j], u[j], acc_q);
            }
          } else {
            float g[8], be[8];
            unpack8(gv[pr], g);
            unpack8(bv[pr], be);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float y = fmaf((u[j] - mu1) * r1, g[j], be[j]);
              zacc[j] = fmaf(w2[pr][j], silu_f(y), zacc[j]);
            }
          }
        }
        if (PHASE == 2) {
          if (L.z_f32) {
            float* zp = reinterpret_cast<float*>(L.z) + off;
            reinterpret_cast<float4*>(zp)[0] = make_float4(zacc[0], zacc[1], zacc[2], zacc[3]);
            reinterpret_cast<float4*>(zp)[1] = make_float4(zacc[4], zacc[5], zacc[6], zacc[7]);
          } else {
            uint4 u4;
            u4.x = Cvt<T>::pack2(zacc[0], zacc[1]); u4.y = Cvt<T>::pack2(zacc[2], zacc[3]);
            u4.z = Cvt<T>::pack2(zacc[4], zacc[5]); u4.w = Cvt<T>::pack2(zacc[6], zacc[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<T*>(L.z) + off) = u4;
            unpack8(u4, zacc);  // statistics of the values phase 3 will read
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc_s += zacc[j];
            acc_q = fmaf(zacc[j], zacc[j], acc_q);
          }
        }
      }
    }
    block_reduce2_to_global(acc_s, acc_q, L.stats + b * 4 + (PHASE == 1 ? 0 : 2));
  } else {
    const double n = 1.0 * C * hw;
    const double m = L.stats[b * 4 + 2] / n;
    const double var = L.stats[b * 4 + 3] / n - m * m;
    const float mu2 = static_cast<float>(m);
    const float r2 = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
    float w3[8], b3[8];
    if (active) {
      fload8(L.w3 + ch, w3);
      fload8(L.b3 + ch, b3);
    }
    // optional GroupNorm statistics of the written rows (this thread's 8 channels touch at most two groups)
    __shared__ float gstat[66][2];
    const bool gn = L.gn_ws != nullptr;
    const int g_lo = gn ? (L.gn_col0 + ch) / L.gn_cpg : 0;
    const int g_split = gn ? (g_lo + 1) * L.gn_cpg - (L.gn_col0 + ch) : 8;  // channels [0, g_split) -> g_lo
    const int g_base = gn ? L.gn_col0 / L.gn_cpg : 0;
    float gs0 = 0.f, gq0 = 0.f, gs1 = 0.f, gq1 = 0.f;
    if (gn) {
      for (int i = threadIdx.x; i < 66 * 2; i += kMergeThreads) (&gstat[0][0])[i] = 0.f;
      __syncthreads();
    }
    if (active) {
#pragma unroll 2
      for (int p = p0 + ty; p < p1; p += ny) {
        const long long off = img_off + static_cast<long long>(p) * C + ch;
        // (all loads of the pixel first, then the arithmetic)
        const long long goff = static_cast<long long>(p) * C + ch;
        const long long row = static_cast<long long>(b) * hw + p;
        const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(L.g2) + goff));
        const uint4 b4 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(L.be2) + goff));
        uint4 s4 = make_uint4(0, 0, 0, 0);
        if (L.skip) s4 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(L.skip) + row * L.lds + ch));
        float zz[8];
        if (L.z_f32) {
          const float4 z0 = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(L.z) + off));
          const float4 z1 = __ldcs(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(L.z) + off) + 1);
          zz[0] = z0.x; zz[1] = z0.y; zz[2] = z0.z; zz[3] = z0.w; zz[4] = z1.x; zz[5] = z1.y; zz[6] = z1.z; zz[7] = z1.w;
        } else {
          mload8_cs<T>(reinterpret_cast<const T*>(L.z) + off, zz);
        }
        float g[8], be[8], sk[8];
        {
          const uint4 uu[3] = {g4, b4, s4};
          float* dst3[3] = {g, be, sk};
#pragma unroll
          for (int q3 = 0; q3 < 3; ++q3) {
            const float2 a = Cvt<T>::unpack2(uu[q3].x), b2 = Cvt<T>::unpack2(uu[q3].y), c2 = Cvt<T>::unpack2(uu[q3].z),
                         d2 = Cvt<T>::unpack2(uu[q3].w);
            dst3[q3][0] = a.x; dst3[q3][1] = a.y; dst3[q3][2] = b2.x; dst3[q3][3] = b2.y;
            dst3[q3][4] = c2.x; dst3[q3][5] = c2.y; dst3[q3][6] = d2.x; dst3[q3][7] = d2.y;
          }
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = sk[j] + fmaf(w3[j], silu_f(fmaf((zz[j] - mu2) * r2, g[j], be[j])), b3[j]);
        uint4 u;
        u.x = Cvt<T>::pack2(o[0], o[1]); u.y = Cvt<T>::pack2(o[2], o[3]);
        u.z = Cvt<T>::pack2(o[4], o[5]); u.w = Cvt<T>::pack2(o[6], o[7]);
        if (gn) {
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = Cvt<T>::unpack2(uu[j]);  // the rounded values the consumer will read
            if (2 * j < g_split) { gs0 += f.x; gq0 = fmaf(f.x, f.x, gq0); } else { gs1 += f.x; gq1 = fmaf(f.x, f.x, gq1); }
            if (2 * j + 1 < g_split) { gs0 += f.y; gq0 = fmaf(f.y, f.y, gq0); } else { gs1 += f.y; gq1 = fmaf(f.y, f.y, gq1); }
          }
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<T*>(L.dst) + row * L.ldd + ch) = u;
      }
    }
    if (gn) {
      if (active) {
        atomicAdd(&gstat[g_lo - g_base][0], gs0);
        atomicAdd(&gstat[g_lo - g_base][1], gq0);
        if (g_split < 8) {
          atomicAdd(&gstat[g_lo - g_base + 1][0], gs1);
          atomicAdd(&gstat[g_lo - g_base + 1][1], gq1);
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 66 * 2; i += kMergeThreads) {
        const float val = (&gstat[0][0])[i];
        if (val != 0.f)
          atomicAdd(L.gn_ws + (static_cast<long long>(b) * L.gn_groups + g_base + (i >> 1)) * 2 + (i & 1), val);
      }
    }
  }
}

template <typename T>
static int merge_levels_t(const MergeTable& tb, int total_ctas, int phase, cudaStream_t s) {
  const dim3 grid(total_ctas), block(kMergeThreads);
  if (phase == 1) ES_CUDA(launch_kernel(merge_levels_kernel<T, 1>, grid, block, 0, s, tb));
  else if (phase == 2) ES_CUDA(launch_kernel(merge_levels_kernel<T, 2>, grid, block, 0, s, tb));
  else ES_CUDA(launch_kernel(merge_levels_kernel<T, 3>, grid, block, 0, s, tb));
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es

extern "C" int es_merge_levels(const EsMergeBatch* m, int phase, void* stream) {
  ES_CHECK(m && phase >= 1 && phase <= 3, "es_merge_levels: bad arguments");
  ES_CHECK(m->n_levels >= 1 && m->n_levels <= ES_MERGE_MAX_LEVELS && m->B >= 1 && m->scale,
           "es_merge_levels: bad batch (levels %d, B %d)", m->n_levels, m->B);
  static_assert(sizeof(es::MergeTable) <= 4096, "kernel parameter space");
  es::MergeTable tb;
  memset(&tb, 0, sizeof(tb));
  tb.n_levels = m->n_levels;
  tb.B = m->B;
  tb.scale = m->scale;
  int cta = 0;
  for (int i = 0; i < m->n_levels; ++i) {
    const EsMergeLevel& s = m->levels[i];
    es::MergeLevelK& k = tb.lv[i];
    ES_CHECK(s.C > 0 && s.C % 8 == 0 && s.C / 8 <= es::kMergeThreads && s.hw > 0 && s.stats && s.z,
             "es_merge_levels: level %d: bad shape/workspace (C %d, hw %d)", i, s.C, s.hw);
    if (phase == 3) ES_CHECK(s.dst && s.ldd % 8 == 0 && (!s.skip || s.lds % 8 == 0), "es_merge_levels: level %d: bad dst", i);
    if (phase == 3 && s.gn_ws)
      ES_CHECK(s.gn_cpg >= 8 && s.gn_col0 >= 0 && s.gn_groups > 0 && s.C / s.gn_cpg + 2 <= 66 &&
                   (s.gn_col0 + s.C + s.gn_cpg - 1) / s.gn_cpg <= s.gn_groups,
               "es_merge_levels: level %d: bad GroupNorm slice", i);
    for (int j = 0; j < 6; ++j) {
      ES_CHECK(s.res[j], "es_merge_levels: level %d: residual %d is NULL", i, j);
      k.res[j] = s.res[j];
    }
    k.w1 = s.w1; k.b1 = s.b1; k.w2 = s.w2; k.b2 = s.b2; k.w3 = s.w3; k.b3 = s.b3;
    k.g1 = s.g1; k.be1 = s.be1; k.g2 = s.g2; k.be2 = s.be2;
    k.stats = s.stats; k.z = s.z; k.skip = s.skip; k.dst = s.dst;
    k.gn_ws = phase == 3 ? s.gn_ws : nullptr;
    k.lds = s.lds; k.ldd = s.ldd; k.hw = s.hw; k.C = s.C;
    k.gn_groups = s.gn_groups; k.gn_cpg = s.gn_cpg; k.gn_col0 = s.gn_col0;
    k.z_f32 = s.z_f32;
    k.gain = s.gain;
    const int ny = es::kMergeThreads / (s.C / 8);
    k.rows_per_cta = ny * es::kMergeIters;
    k.chunks = (s.hw + k.rows_per_cta - 1) / k.rows_per_cta;
    k.cta0 = cta;
    cta += k.chunks * m->B;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return m->dtype == ES_DTYPE_BF16 ? es::merge_levels_t<__nv_bfloat16>(tb, cta, phase, st)
                                   : es::merge_levels_t<__half>(tb, cta, phase, st);
}
