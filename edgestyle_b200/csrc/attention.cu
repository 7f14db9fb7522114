// es_attention: flash-style softmax(Q K^T * scale) V on tcgen05/TMEM tiles fed by TMA (sm_100a).
//
//   grid (ceil(nq/128), heads, batch), 192 threads:
//     warp 0     TMA producer: Q once, then one K tile and one V tile (128 keys) per iteration.  Head
//                dim d is its own tensor-map dimension, so the box's columns d..DP-1 are out of bounds and
//                zero-filled by TMA -- d = 40/80/160 need no padding in memory.
//     warp 1     TMEM allocator + single-thread MMA issuer:
//                  S[128x128]  = Q[128xd] K[128xd]^T      (both K-major, SWIZZLE_128B)    -> TMEM cols 0..127
//                  O[128xdN]  += P[128x128] V[128xdN]     (P K-major from smem, V MN-major) -> TMEM cols 128..
//     warps 2-5  softmax: thread r owns query row r (TMEM lane r): online max/sum in fp32 with exp2,
//                P written to smem as 16-bit in the canonical K-major SWIZZLE_128B layout, O rescaled in TMEM
//                only when some row's max moved, final O / l stored with 16 B vectors.
// Replaces F.scaled_dot_product_attention in diffusers' AttnProcessor2_0 (see include/edgestyle_b200.h).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace es {

constexpr int kAttThreads = 192;
constexpr int kAtomBytes = 128 * 128;  // 128 rows x 128 B
constexpr int kKV = 64;                // keys per K/V tile

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA pipe (Cody-Waite split + degree-4 polynomial, relative error < 5e-5 -- below the 16-bit P it feeds):
// the softmax is bound by the MUFU pipe (one ex2 per score), so a fraction of the exponentials is taken off it.
#ifndef ES_ATT_POLY_EVERY
#define ES_ATT_POLY_EVERY 0  // 0: all exponentials on MUFU; n: every n-th score of a row uses the polynomial
#endif
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.f);
  const float magic = 12582912.f;  // 1.5 * 2^23: adding it rounds x to the nearest integer in the low mantissa bits
  const float xr = x + magic;
  const float r = x - (xr - magic);  // [-0.5, 0.5]
  float p = fmaf(r, 0.0096181291f, 0.0555041087f);
  p = fmaf(p, r, 0.2402265070f);
  p = fmaf(p, r, 0.6931471806f);
  p = fmaf(p, r, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}

// Optional event trace (clock64 stamps of one CTA) for pipeline analysis: build with -DES_ATT_TRACE.
#ifdef ES_ATT_TRACE
__device__ long long g_att_trace[4][64];
#define ATT_TRACE(role, slot)                                                            \
  do {                                                                                   \
    if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (slot) < 64) g_att_trace[role][slot] = clock64(); \
  } while (0)
#else
#define ATT_TRACE(role, slot) do {} while (0)
#endif

struct AttParams {
  int d, dN, nq, nkv;
  float scale_log2;
  void* out;
  long long ldo, bso;
  uint32_t tmem_cols;  // attention2_kernel: TMEM columns to allocate
  int o_stride;        // attention2_kernel: TMEM columns between the O accumulators of two query tiles
};

template <typename T, int NA>
__global__ void __launch_bounds__(kAttThreads, NA == 1 ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  // TMEM: S double-buffered (2 x 64 fp32 columns) + O (dN columns)
  constexpr uint32_t kTmemCols = NA <= 2 ? 256 : 512;
  constexpr int kOCol = 128;
  constexpr int kKVAtom = kKV * 128;  // bytes of one 64-column atom of a K / V tile (kKV rows x 128 B)
  // No static shared memory at all: the dynamic region then starts at the CTA's (1024-byte aligned) window base, so
  // SWIZZLE_128B needs no alignment slack and two CTAs of the d <= 64 variant (80 KB each) fit one SM.
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                              // NA atoms x [128 rows x 128 B]
  uint8_t* sK = sQ + NA * kAtomBytes;              // 2 stages x NA atoms x [64 rows x 128 B]
  uint8_t* sV = sK + 2 * NA * kKVAtom;             // 2 stages x NA atoms x [64 rows x 128 B]
  uint8_t* sP = sV + 2 * NA * kKVAtom;             // 2 buffers x [128 rows x 128 B] (64 keys = one 128 B row)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kAtomBytes);
  uint64_t& q_full = bars[0];
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t& o_done = bars[13];
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 14);

  pdl_launch_dependents();
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("edgestyle_b200: attention smem base not 1024-byte aligned\n");
    __trap();
  }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.nkv + kKV - 1) / kKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
    }
    mbar_init(&o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // PDL: everything above overlaps the previous kernel's tail; its outputs may only be read after this point.
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer: Q once, then 2-stage K and V rings ==========================
    if (lane == 0) {
      mbar_expect_tx(&q_full, NA * kAtomBytes);
#pragma unroll
      for (int a = 0; a < NA; ++a) tma_load_4d(sQ + a * kAtomBytes, &tmQ, &q_full, a * 64, head, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], NA * kKVAtom);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_4d(sK + (s * NA + a) * kKVAtom, &tmK, &k_full[s], a * 64, head, j * kKV, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], NA * kKVAtom);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_4d(sV + (s * NA + a) * kKVAtom, &tmV, &v_full[s], a * 64, head, j * kKV, b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================================================
    // S(j+2) = Q K(j+2)^T is issued as soon as softmax(j) has consumed S[j & 1], i.e. two tiles ahead of the PV that
    // needs it: the softmax warps never wait for a QK^T, and PV(j) retires while softmax(j+1) runs.
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_f16(128, kKV, Cvt<T>::kFmt, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, p.dN, Cvt<T>::kFmt, 0, 1);
      const int kq = (p.d + 15) / 16;  // MMAs along the head dim
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      mbar_wait(&q_full, 0);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        for (int k = 0; k < kq; ++k) {
          const uint64_t ad = smem_desc_sw128(aQ + (k >> 2) * kAtomBytes, 16, 1024) + 2 * (k & 3);
          const uint64_t bd = smem_desc_sw128(aK + (s * NA + (k >> 2)) * kKVAtom, 16, 1024) + 2 * (k & 3);
          umma_f16(tmem_base + s * kKV, ad, bd, idesc_qk, k != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      };
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ATT_TRACE(0, 4 * j + 1);
        mbar_wait(&p_full[s], ph);  // P(j) in smem and S[s] consumed
        ATT_TRACE(0, 4 * j + 2);
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kKV / 16; ++k) {
          const uint64_t ad = smem_desc_sw128(aP + s * kAtomBytes, 16, 1024) + 2 * k;
          // V: MN-major, 16 keys = 2048 B along K; N atoms (64 of d) kKVAtom apart
          const uint64_t bd = smem_desc_sw128(aV + s * NA * kKVAtom + k * 2048, kKVAtom, 1024);
          umma_f16(tmem_base + kOCol, ad, bd, idesc_pv, (j | k) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(&o_done);
        ATT_TRACE(0, 4 * j + 3);
        if (j + 2 < n_tiles) issue_qk(j + 2);
        ATT_TRACE(0, 4 * j + 0);
      }
    }
  } else {
    // ================================ softmax: thread r owns query row r (TMEM lane r) ==========================
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    float m_run = -INFINITY;  // row max estimate used as the softmax shift, in scaled (log2) units
    float l_run = 0.f;        // running row sum of P (fp32)
    const float sl2 = p.scale_log2;
    // Softmax is shift-invariant, so the shift only has to stay within 2^kSlack of the true running max
    // (P <= 2^kSlack fits 16-bit floats).  Tiles after the first are therefore processed in ONE pass with
    // the previous shift; a warp re-does its 32 rows (and rescales O) only when some row's max grew by
    // more than kSlack -- rare once the first tiles have been seen.
    constexpr float kSlack = 8.0f;

    // writes P for columns [c, c+32) of this thread's row from raw scores v, with shift -neg_m
    auto emit_p = [&](const uint32_t (&v)[32], int c, float neg_m, int kv_valid, bool full_tile, float& sum, int pbuf) {
      uint32_t pk[16];
      float s0 = 0.f, s1 = 0.f;
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a0 = fmaf(__uint_as_float(v[i]), sl2, neg_m), a1 = fmaf(__uint_as_float(v[i + 1]), sl2, neg_m);
          const float p0 = ex2_approx(a0);
          const float p1 = (ES_ATT_POLY_EVERY > 0 && ((i + 1) % ES_ATT_POLY_EVERY) == ES_ATT_POLY_EVERY - 1)
                               ? ex2_poly(a1) : ex2_approx(a1);
          s0 += p0;
          s1 += p1;
          pk[i >> 1] = Cvt<T>::pack2(p0, p1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = (c + i < kv_valid) ? ex2_approx(fmaf(__uint_as_float(v[i]), sl2, neg_m)) : 0.f;
          const float p1 = (c + i + 1 < kv_valid) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, neg_m)) : 0.f;
          s0 += p0;
          s1 += p1;
          pk[i >> 1] = Cvt<T>::pack2(p0, p1);
        }
      }
      sum += s0 + s1;
      // canonical K-major SWIZZLE_128B: row r at r*128 B, 16 B chunk index XOR (r & 7)
      uint8_t* prow = sP + pbuf * kAtomBytes + r * 128;
      const int cc0 = c >> 3;  // first 16 B chunk of this 32-column group
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int chunk = (cc0 + q4) ^ (r & 7);
        *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
      }
    };
    auto tile_max = [&](const uint32_t (&v)[32], int c, int kv_valid, bool full_tile, float mx) {
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = fmax3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      return mx;
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kv_valid = min(kKV, p.nkv - j * kKV);
      const bool full_tile = kv_valid == kKV;
      const uint32_t t_s = t_row + s * kKV;
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 0);
      mbar_wait(&s_full[s], ph);
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 1);
      // P[s] is free: QK(j) completes after PV(j-2) on the in-order tensor pipe.  Only the (rare) O rescale below
      // has to wait for PV(j-1).
      tc_fence_after();
      float mx = -INFINITY;
      uint32_t va[32], vb[32];
      if (j == 0) {
        // first tile: a max pass to establish the shift
        tmem_ld_x32(t_s, va);
        tmem_ld_wait();
        tmem_ld_x32(t_s + 32, vb);
        mx = tile_max(va, 0, kv_valid, full_tile, mx);
        tmem_ld_wait();
        mx = tile_max(vb, 32, kv_valid, full_tile, mx);
        m_run = mx * sl2;
      }
      float sum = 0.f;
      for (int attempt = 0; attempt < 2; ++attempt) {
        const float neg_m = -m_run;
        sum = 0.f;
        tmem_ld_x32(t_s, va);
        tmem_ld_wait();
        tmem_ld_x32(t_s + 32, vb);  // in flight while the first half is exponentiated
        if (j > 0 && attempt == 0) mx = tile_max(va, 0, kv_valid, full_tile, mx);
        emit_p(va, 0, neg_m, kv_valid, full_tile, sum, s);
        tmem_ld_wait();
        if (j > 0 && attempt == 0) mx = tile_max(vb, 32, kv_valid, full_tile, mx);
        emit_p(vb, 32, neg_m, kv_valid, full_tile, sum, s);
        if (j == 0 || attempt == 1) break;
        const float m_tile = mx * sl2;
        if (!__any_sync(0xffffffffu, m_tile > m_run + kSlack)) break;
        // rare path: raise the shift of the rows that need it, rescale their O, redo P
        mbar_wait(&o_done, (j - 1) & 1);  // PV(j-1) must have retired before O is touched
        tc_fence_after();
        const float m_new = fmaxf(m_run, m_tile);
        const float alpha = ex2_approx(m_run - m_new);
#pragma unroll 1
        for (int c = 0; c < p.dN; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(t_row + kOCol + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(t_row + kOCol + c, o);
        }
        tmem_st_wait();
        l_run *= alpha;
        m_run = m_new;
      }
      l_run += sum;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(&p_full[s]);
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 3);
    }
    // epilogue: O / l
    mbar_wait(&o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const bool row_ok = (q0 + r) < p.nq;
    T* optr = reinterpret_cast<T*>(p.out) + static_cast<long long>(b) * p.bso +
              static_cast<long long>(q0 + r) * p.ldo + head * p.d;
#pragma unroll 1
    for (int c = 0; c < p.dN; c += 16) {
      uint32_t o[16];
      tmem_ld_x16(t_row + kOCol + c, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c + h * 8 < p.d) {
            uint4 u;
            u.x = Cvt<T>::pack2(__uint_as_float(o[h * 8 + 0]) * inv_l, __uint_as_float(o[h * 8 + 1]) * inv_l);
            u.y = Cvt<T>::pack2(__uint_as_float(o[h * 8 + 2]) * inv_l, __uint_as_float(o[h * 8 + 3]) * inv_l);
            u.z = Cvt<T>::pack2(__uint_as_float(o[h * 8 + 4]) * inv_l, __uint_as_float(o[h * 8 + 5]) * inv_l);
            u.w = Cvt<T>::pack2(__uint_as_float(o[h * 8 + 6]) * inv_l, __uint_as_float(o[h * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(optr + c + h * 8) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <typename T, int NA>
static int launch_attention(const EsAttention* a, cudaStream_t stream) {
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[4] = {64u, 1u, 128u, 1u};
  {
    uint64_t dims[4] = {(uint64_t)a->d, (uint64_t)a->heads, (uint64_t)a->nq, (uint64_t)a->batch};
    uint64_t str[4] = {0, (uint64_t)a->d * 2, (uint64_t)a->ldq * 2, (uint64_t)a->bsq * 2};
    if (encode_tmap_16b(&tmQ, a->q, 4, dims, str, box)) return -3;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->d, (uint64_t)a->heads, (uint64_t)a->nkv, (uint64_t)a->batch};
    uint64_t str[4] = {0, (uint64_t)a->d * 2, (uint64_t)a->ldk * 2, (uint64_t)a->bsk * 2};
    const uint32_t boxkv[4] = {64u, 1u, static_cast<uint32_t>(kKV), 1u};
    if (encode_tmap_16b(&tmK, a->k, 4, dims, str, boxkv)) return -3;
    uint64_t strv[4] = {0, (uint64_t)a->d * 2, (uint64_t)a->ldv * 2, (uint64_t)a->bsv * 2};
    if (encode_tmap_16b(&tmV, a->v, 4, dims, strv, boxkv)) return -3;
  }
  AttParams p;
  p.d = a->d;
  p.dN = ((a->d + 15) / 16) * 16;
  p.nq = a->nq;
  p.nkv = a->nkv;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = a->out;
  p.ldo = a->ldo;
  p.bso = a->bso;
  const size_t smem = static_cast<size_t>(NA + 2) * kAtomBytes + 4 * NA * kKV * 128 + 128;  // Q + 2 P + 2x(K,V) + barriers
  auto kern = attention_kernel<T, NA>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  dim3 grid((a->nq + 127) / 128, a->heads, a->batch);
  ES_CUDA(launch_kernel(kern, dim3(grid), dim3(kAttThreads), smem, stream, tmQ, tmK, tmV, p));
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es
#include "attention2.cuh"
namespace es {

// 0 = second-generation kernel (attention2.cuh), 1 = the first one (kept for A/B measurements: ES_ATT_V1=1)
static int att_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ES_ATT_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

template <typename T>
static int attention2_dispatch(const EsAttention* a, cudaStream_t s) {
  // two query tiles per CTA when that still gives every SM work (K/V tiles are then loaded once for both)
  const long long ctas2 = static_cast<long long>((a->nq + 255) / 256) * a->heads * a->batch;
  const bool two = a->nq > 128 && ctas2 >= 148;
  if (a->d <= 64) return two ? launch_attention2<T, 1, 2>(a, s) : launch_attention2<T, 1, 1>(a, s);
  if (a->d <= 128) return two ? launch_attention2<T, 2, 2>(a, s) : launch_attention2<T, 2, 1>(a, s);
  return two ? launch_attention2<T, 3, 2>(a, s) : launch_attention2<T, 3, 1>(a, s);
}

template <typename T>
static int attention_dispatch(const EsAttention* a, cudaStream_t s) {
  ES_CHECK(a->d % 8 == 0 && a->d >= 8 && a->d <= 192, "es_attention: head dim %d unsupported (multiple of 8, <= 192)", a->d);
  ES_CHECK(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 && a->bsq % 8 == 0 &&
               a->bsk % 8 == 0 && a->bsv % 8 == 0 && a->bso % 8 == 0,
           "es_attention: pitches must be multiples of 8 elements");
  ES_CHECK(a->nq > 0 && a->nkv > 0 && a->batch > 0 && a->heads > 0, "es_attention: empty problem");
  if (!att_v1()) return attention2_dispatch<T>(a, s);
  if (a->d <= 64) return launch_attention<T, 1>(a, s);
  if (a->d <= 128) return launch_attention<T, 2>(a, s);
  return launch_attention<T, 3>(a, s);
}

}  // namespace es

#ifdef ES_ATT_TRACE
extern "C" int es_attention_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, es::g_att_trace, sizeof(es::g_att_trace)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int es_attention(const EsAttention* a, void* stream) {
  if (!a) {
    es::set_error("es_attention: null descriptor");
    return -1;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == ES_DTYPE_BF16) return es::attention_dispatch<__nv_bfloat16>(a, s);
  return es::attention_dispatch<__half>(a, s);
}
