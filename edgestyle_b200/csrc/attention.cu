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
#include "common.h"
#include "ptx.cuh"

namespace es {

constexpr int kAttThreads = 192;
constexpr int kAtomBytes = 128 * 128;  // 128 rows x 128 B

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
};

template <typename T, int NA>
__global__ void __launch_bounds__(kAttThreads, NA == 1 ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  constexpr uint32_t kTmemCols = NA <= 2 ? 256 : 512;
  constexpr int kOCol = 128;
  // No static shared memory at all: the dynamic region then starts at the CTA's (1024-byte aligned) window base, so
  // SWIZZLE_128B needs no alignment slack and two CTAs of the d<=64 variant (7 x 16 KB + barriers each) fit one SM.
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NA * kAtomBytes;
  uint8_t* sV = sK + NA * kAtomBytes;
  uint8_t* sP = sV + NA * kAtomBytes;  // 2 buffers x 2 atoms: softmax(j+1) writes one while PV(j) reads the other
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kAtomBytes);
  uint64_t& q_full = bars[0];
  uint64_t& k_full = bars[1];
  uint64_t& v_full = bars[2];
  uint64_t& k_empty = bars[3];
  uint64_t& v_empty = bars[4];
  uint64_t& s_full = bars[5];
  uint64_t& p_full = bars[6];
  uint64_t& o_done = bars[7];
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 8);

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
  const int n_tiles = (p.nkv + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full, 1);
    mbar_init(&k_full, 1);
    mbar_init(&v_full, 1);
    mbar_init(&k_empty, 1);
    mbar_init(&v_empty, 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_full, 128);
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
  // PDL: everything above (barrier init, TMEM alloc, descriptor prefetch) overlaps the previous kernel's tail;
  // global memory written by it may only be touched after this point.
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&q_full, NA * kAtomBytes);
#pragma unroll
      for (int a = 0; a < NA; ++a) tma_load_4d(sQ + a * kAtomBytes, &tmQ, &q_full, a * 64, head, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(&k_empty, ph ^ 1);
        mbar_expect_tx(&k_full, NA * kAtomBytes);
#pragma unroll
        for (int a = 0; a < NA; ++a) tma_load_4d(sK + a * kAtomBytes, &tmK, &k_full, a * 64, head, j * 128, b);
        mbar_wait(&v_empty, ph ^ 1);
        mbar_expect_tx(&v_full, NA * kAtomBytes);
#pragma unroll
        for (int a = 0; a < NA; ++a) tma_load_4d(sV + a * kAtomBytes, &tmV, &v_full, a * 64, head, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_f16(128, 128, Cvt<T>::kFmt, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, p.dN, Cvt<T>::kFmt, 0, 1);
      const int kq = (p.d + 15) / 16;  // MMAs along the head dim
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      mbar_wait(&q_full, 0);
      auto issue_qk = [&](int j) {
        mbar_wait(&k_full, j & 1);
        tc_fence_after();
        for (int k = 0; k < kq; ++k) {
          const uint64_t ad = smem_desc_sw128(aQ + (k >> 2) * kAtomBytes, 16, 1024) + 2 * (k & 3);
          const uint64_t bd = smem_desc_sw128(aK + (k >> 2) * kAtomBytes, 16, 1024) + 2 * (k & 3);
          umma_f16(tmem_base, ad, bd, idesc_qk, k != 0);
        }
        umma_commit(&k_empty);
        umma_commit(&s_full);
      };
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        ATT_TRACE(0, 4 * j + 1);
        mbar_wait(&p_full, ph);  // P(j) in smem and S(j) consumed
        ATT_TRACE(0, 4 * j + 2);
        // S is free again: start the next tile's scores first so its softmax can begin while PV(j) runs
        if (j + 1 < n_tiles) issue_qk(j + 1);
        ATT_TRACE(0, 4 * j + 0);
        mbar_wait(&v_full, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 128 keys = 8 x K16
          const uint64_t ad = smem_desc_sw128(aP + ((j & 1) * 2 + (k >> 2)) * kAtomBytes, 16, 1024) + 2 * (k & 3);
          // V: MN-major, 16 keys = 2048 B along K; N atoms (64 of d) kAtomBytes apart
          const uint64_t bd = smem_desc_sw128(aV + k * 2048, kAtomBytes, 1024);
          umma_f16(tmem_base + kOCol, ad, bd, idesc_pv, (j | k) != 0);
        }
        umma_commit(&v_empty);
        umma_commit(&o_done);
        ATT_TRACE(0, 4 * j + 3);
      }
    }
  } else {
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
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, neg_m));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sl2, neg_m));
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
      uint8_t* prow = sP + (pbuf * 2 + (c >> 6)) * kAtomBytes + r * 128;
      const int cc0 = (c & 63) >> 3;  // first 16 B chunk of this 32-column group inside the atom
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
      const uint32_t ph = j & 1;
      const int kv_valid = min(128, p.nkv - j * 128);
      const bool full_tile = kv_valid == 128;
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 0);
      mbar_wait(&s_full, ph);
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 1);
      // P is double-buffered and QK(j) completes after PV(j-2) on the in-order tensor pipe, so P[j & 1] is free here;
      // only the (rare) O rescale below has to wait for PV(j-1)
      tc_fence_after();
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 2);
      float mx = -INFINITY;
      uint32_t va[32], vb[32];
      if (j == 0) {
        // first tile: a max pass to establish the shift (TMEM loads software-pipelined against the math)
        tmem_ld_x32(t_row, va);
        tmem_ld_wait();
        tmem_ld_x32(t_row + 32, vb);
        mx = tile_max(va, 0, kv_valid, full_tile, mx);
        tmem_ld_wait();
        tmem_ld_x32(t_row + 64, va);
        mx = tile_max(vb, 32, kv_valid, full_tile, mx);
        tmem_ld_wait();
        tmem_ld_x32(t_row + 96, vb);
        mx = tile_max(va, 64, kv_valid, full_tile, mx);
        tmem_ld_wait();
        mx = tile_max(vb, 96, kv_valid, full_tile, mx);
        m_run = mx * sl2;
      }
      bool redo = false;
      float sum = 0.f;
      for (int attempt = 0; attempt < 2; ++attempt) {
        const float neg_m = -m_run;
        sum = 0.f;
        tmem_ld_x32(t_row, va);
        tmem_ld_wait();
        tmem_ld_x32(t_row + 32, vb);
        if (j > 0 && attempt == 0) mx = tile_max(va, 0, kv_valid, full_tile, mx);
        emit_p(va, 0, neg_m, kv_valid, full_tile, sum, j & 1);
        tmem_ld_wait();
        tmem_ld_x32(t_row + 64, va);
        if (j > 0 && attempt == 0) mx = tile_max(vb, 32, kv_valid, full_tile, mx);
        emit_p(vb, 32, neg_m, kv_valid, full_tile, sum, j & 1);
        tmem_ld_wait();
        tmem_ld_x32(t_row + 96, vb);
        if (j > 0 && attempt == 0) mx = tile_max(va, 64, kv_valid, full_tile, mx);
        emit_p(va, 64, neg_m, kv_valid, full_tile, sum, j & 1);
        tmem_ld_wait();
        if (j > 0 && attempt == 0) mx = tile_max(vb, 96, kv_valid, full_tile, mx);
        emit_p(vb, 96, neg_m, kv_valid, full_tile, sum, j & 1);
        if (j == 0 || attempt == 1) break;
        const float m_tile = mx * sl2;
        redo = __any_sync(0xffffffffu, m_tile > m_run + kSlack);
        if (!redo) break;
        // rare path: raise the shift of the rows that need it, rescale their O / row-sum columns, redo P
        mbar_wait(&o_done, ph ^ 1);  // PV(j-1) must have retired before O is touched
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
      mbar_arrive(&p_full);
      if (threadIdx.x == 64) ATT_TRACE(1, 4 * j + 3);
    }
    // epilogue: O / l   (l = row sum accumulated by the tensor core in the columns right after O)
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
    if (encode_tmap_16b(&tmK, a->k, 4, dims, str, box)) return -3;
    uint64_t strv[4] = {0, (uint64_t)a->d * 2, (uint64_t)a->ldv * 2, (uint64_t)a->bsv * 2};
    if (encode_tmap_16b(&tmV, a->v, 4, dims, strv, box)) return -3;
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
  const size_t smem = static_cast<size_t>(3 * NA + 4) * kAtomBytes + 128;  // + barriers; no alignment slack (see kernel)
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

template <typename T>
static int attention_dispatch(const EsAttention* a, cudaStream_t s) {
  ES_CHECK(a->d % 8 == 0 && a->d >= 8 && a->d <= 192, "es_attention: head dim %d unsupported (multiple of 8, <= 192)", a->d);
  ES_CHECK(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0 && a->bsq % 8 == 0 &&
               a->bsk % 8 == 0 && a->bsv % 8 == 0 && a->bso % 8 == 0,
           "es_attention: pitches must be multiples of 8 elements");
  ES_CHECK(a->nq > 0 && a->nkv > 0 && a->batch > 0 && a->heads > 0, "es_attention: empty problem");
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
