// es_gemm, CTA-pair variant: persistent tcgen05 `cta_group::2` implicit GEMM with 256 x 320 tiles.
//
// Why: at 128 x 160 tiles the long-K convolutions of the 64x64 / 32x32 levels are bound by L2 -> SM operand traffic
// (every CTA streams 16 KB of A and 20 KB of B per 64-wide K block: 829 MB for the M = 32768, N = 320, K = 2880 conv
// against a measured ~12 TB/s), not by the tensor pipe.  Two SMs of a TPC that share one 320-row weight tile (each
// loads half of it) and keep a 128 x 320 accumulator each in TMEM halve that traffic; N = 320 divides every wide
// layer of SD1.5 (320, 640, 960, 1280, 1920, 2560, 3840, 5120, 10240).
//
//   cluster    : 2 CTAs along M (rank r owns tile rows [128 r, 128 r + 128) of the pair's 256), 1 CTA / SM, persistent
//                over work units u = (m pair, n tile, k split), u += number of clusters.
//   warp 0     : TMA producer in BOTH CTAs: own A box (16 KB) + own half of B (two boxes of 80 rows: rows
//                [160 j + 80 r, +80) of the tile for the two UMMA halves j); completion bytes go to the LEADER's full
//                barrier (`.cta_group::2` TMA, barrier address via mapa).
//   warp 1     : TMEM allocation (both CTAs, 512 columns), and in the leader the single-thread issue of
//                tcgen05.mma.cta_group::2 (M = 256, N = 160, twice per K step); tcgen05.commit multicasts the
//                "slot free" / "accumulator ready" arrivals to the barriers of both CTAs.
//   warps 2..9 : epilogue, two warps per TMEM lane quarter (columns [0,160) and [160,320)): per-column vector
//                (bias + time-embedding row) prefetched while the mainloop runs, pipelined tcgen05.ld, GEGLU /
//                residual / GroupNorm statistics, swizzled smem panels in a DEDICATED region, TMA store.  The smem
//                ring is independent of the epilogue, so the producer prefetches the next unit during the epilogue.
//   split-K    : as in the single-CTA kernel (partials in the caller's workspace, last arriver runs the epilogue).
#pragma once

namespace es {

#ifdef ES_GEMM_TRACE
#define PAIR_TRACE(slot)                                              \
  do {                                                                \
    if (blockIdx.x == 0) g_gemm_trace[slot] = clock64();              \
  } while (0)
#else
#define PAIR_TRACE(slot) do {} while (0)
#endif

constexpr int kPairN = 320;
constexpr int kPairNH = 160;                          // columns per UMMA
constexpr int kPairThreads = 320;
constexpr int kPairStages = 3;
constexpr int kPairABytes = kBlockM * kBlockK * 2;     // 16 KB
constexpr int kPairBHalf = (kPairNH / 2) * kBlockK * 2;  // 80 rows x 128 B
constexpr int kPairBBytes = 2 * kPairBHalf;            // 20 KB: this CTA's half of the 320-row weight tile
constexpr int kPairStageBytes = kPairABytes + kPairBBytes;
constexpr int kPairPanelBytes = 5 * 16384;             // 128 x 320 outputs as five 64-column panels
constexpr int kPairVecBytes = 4 * kPairN * 4;
constexpr int kPairGnBytes = 4 * kGnSlots * 2 * 4;    // fused GroupNorm statistics: [4 images][kGnSlots][2]
constexpr int kPairSmem = kPairStages * kPairStageBytes + kPairPanelBytes + kPairVecBytes + kPairGnBytes + 1024;
constexpr uint32_t kPairTmemCols = 512;

struct PairUnit {
  int x0, y0, i0, x_end, b_noff, b2_noff, n0, tn, tile_m, z, kb_begin, kb_end;
};

__device__ __forceinline__ PairUnit pair_decode(const GemmKParams& p, int u, int m_pairs, int n_tiles, int rank) {
  PairUnit d;
  const int pm = u % m_pairs;
  const int rest = u / m_pairs;
  d.tn = rest % n_tiles;
  d.z = rest / n_tiles;
  d.tile_m = 2 * pm + rank;
  d.n0 = d.tn * kPairN;
  if (p.flat) {
    int g = 0;
#pragma unroll
    for (int s = 1; s < ES_MAX_SEG; ++s)
      if (s < p.nseg && d.tile_m >= p.seg_tile_start[s]) g = s;
    d.x0 = p.seg_row_start[g] + (d.tile_m - p.seg_tile_start[g]) * kBlockM;
    d.x_end = p.seg_row_start[g + 1];
    d.y0 = 0;
    d.i0 = 0;
    d.b_noff = p.seg_b_noff[g];
    d.b2_noff = p.seg_b2_noff[g];
  } else {
    const int tx = d.tile_m % p.tiles_x;
    const int ty = (d.tile_m / p.tiles_x) % p.tiles_y;
    const int tnn = d.tile_m / (p.tiles_x * p.tiles_y);
    d.x0 = tx * p.bw;
    d.y0 = ty * p.bh;
    d.i0 = tnn * p.bn;
    d.x_end = p.W;
    int g = 0;  // image segments (conv LoRA)
#pragma unroll
    for (int s = 1; s < ES_MAX_SEG; ++s)
      if (s < p.nseg && d.i0 >= p.seg_row_start[s]) g = s;
    d.b_noff = p.seg_b_noff[g];
    d.b2_noff = p.seg_b2_noff[g];
  }
  const int kb1 = p.taps * p.kblocks1;
  const int kb_total = kb1 + ((p.kblocks2 > 0 && d.b2_noff >= 0) ? p.kblocks2 : 0);
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  d.kb_begin = d.z * kb_per;
  d.kb_end = min(kb_total, d.kb_begin + kb_per);
  return d;
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmOp,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmRp,
                 const GemmKParams p, const int n_units, const int m_pairs, const int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed on the SHARED-window address: going through uintptr_t would make
  // every later access a generic LD / ST (64-bit address arithmetic, no LDS / STS) -- measured in the epilogues
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* panels = smem + kPairStages * kPairStageBytes;
  float* vec_s = reinterpret_cast<float*>(panels + kPairPanelBytes);  // [4 lane quarters][320]
  float* gstat_s = vec_s + 4 * kPairN;
  __shared__ __align__(8) uint64_t full_bar[kPairStages];
  __shared__ __align__(8) uint64_t empty_bar[kPairStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;   // MMA -> epilogue (both CTAs, multicast commit)
  __shared__ __align__(8) uint64_t tmem_empty_bar;  // epilogue warps of BOTH CTAs -> the leader's MMA issuer
  __shared__ __align__(8) uint64_t res_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int splitk_last;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int kb1 = p.taps * p.kblocks1;
  pdl_launch_dependents();
  if (warp == 2 && lane == 0) l2_prefetch_slice(p.prefetch, p.prefetch_bytes, blockIdx.x, gridDim.x);
  if (threadIdx.x == 0) PAIR_TRACE(0);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kblocks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, 16);  // one arrival per epilogue warp of both CTAs
    mbar_init(&res_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(&tmem_base_smem, kPairTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // PDL: global memory written by the previous kernel may be touched from here on
  if (threadIdx.x == 0) PAIR_TRACE(1);

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if (elect_one()) {
      const uint32_t full0 = mapa_u32(&full_bar[0], 0);  // the leader's full barriers
      int it = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters) {
        const PairUnit d = pair_decode(p, u, m_pairs, n_tiles, rank);
        for (int kb = d.kb_begin; kb < d.kb_end; ++kb, ++it) {
          const int s = it % kPairStages;
          const uint32_t ph = (it / kPairStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * kPairStageBytes;
          uint8_t* sb = sa + kPairABytes;
          if (leader) mbar_expect_tx(&full_bar[s], 2 * kPairStageBytes);
          const uint32_t fb = full0 + s * 8;
          if (kb < kb1) {
            const int tap = kb / p.kblocks1;
            const int cb = kb - tap * p.kblocks1;
            int dx = 0, dy = 0;
            if (p.taps == 9) {
              dy = tap / 3 - 1;
              dx = tap % 3 - 1;
            }
            tma_load_4d_pair(sa, &tmA, fb, cb * kBlockK, d.x0 + dx, d.y0 + dy, d.i0);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_3d_pair(sb + j * kPairBHalf, &tmB, fb, cb * kBlockK, tap,
                               d.b_noff + d.n0 + j * kPairNH + static_cast<int>(rank) * (kPairNH / 2));
          } else {
            const int cb = kb - kb1;
            tma_load_4d_pair(sa, &tmA2, fb, cb * kBlockK, d.x0, d.y0, d.i0);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_3d_pair(sb + j * kPairBHalf, &tmB2, fb, cb * kBlockK, 0,
                               d.b2_noff + d.n0 + j * kPairNH + static_cast<int>(rank) * (kPairNH / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===========================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(2 * kBlockM, kPairNH, Cvt<T>::kFmt, 0, 0);
      const uint32_t peer_full = mapa_u32(&tmem_full_bar, 1);
      int it = 0, ui = 0;
      for (int u = cluster_id; u < n_units; u += n_clusters, ++ui) {
        const PairUnit d = pair_decode(p, u, m_pairs, n_tiles, rank);
        mbar_wait(&tmem_empty_bar, (ui & 1) ^ 1);  // both CTAs' epilogues have drained the accumulator
        tc_fence_after();
        for (int kb = d.kb_begin; kb < d.kb_end; ++kb, ++it) {
          const int s = it % kPairStages;
          const uint32_t ph = (it / kPairStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (it == 0) PAIR_TRACE(2);
          if (it == 8) PAIR_TRACE(7);
          if (it == 24) PAIR_TRACE(8);
          const uint32_t sa = smem_u32(smem + s * kPairStageBytes);
          const uint32_t sb = sa + kPairABytes;
          const uint64_t adesc = smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc0 = smem_desc_sw128(sb, 16, 1024);
          const uint64_t bdesc1 = smem_desc_sw128(sb + kPairBHalf, 16, 1024);
          const uint32_t first = (kb != d.kb_begin);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_f16_pair(tmem_base, adesc + 2 * k, bdesc0 + 2 * k, idesc, first | k);
            umma_f16_pair(tmem_base + kPairNH, adesc + 2 * k, bdesc1 + 2 * k, idesc, first | k);
          }
          umma_commit_pair(&empty_bar[s]);
        }
        if (ui == 0) PAIR_TRACE(3);
        if (d.kb_end > d.kb_begin) {
          umma_commit_pair(&tmem_full_bar);
        } else {  // an empty K range (split-K remainder): nothing to wait for
          mbar_arrive(&tmem_full_bar);
          mbar_arrive_cluster(peer_full);
        }
      }
    }
  } else {
    // =============================== epilogue (8 warps) =====================================
    const int q = warp & 3;             // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;   // accumulator columns [160 * half, +160)
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;    // 0..255
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * kPairNH;
    const uint32_t lead_empty = mapa_u32(&tmem_empty_bar, 0);
    const bool geglu = p.act == ES_ACT_GEGLU;
    const bool has_res = p.residual != nullptr;
    const int n_tile_out = geglu ? kPairNH : kPairN;
    const int full_panels = n_tile_out / 64;
    const int rem = n_tile_out % 64;
    const bool gn_panel = p.gn_ws && !geglu && p.gn_cpg >= 8 && kPairN / p.gn_cpg + 2 <= kGnSlots;
    int ui = 0, res_it = 0;
    for (int u = cluster_id; u < n_units; u += n_clusters, ++ui) {
      const PairUnit d = pair_decode(p, u, m_pairs, n_tiles, rank);
      const bool has_work = d.kb_end > d.kb_begin;
      const int x0 = d.x0, y0 = d.y0, i0 = d.i0, n0 = d.n0, b_noff = d.b_noff;
      const int oc0 = geglu ? d.tn * kPairNH : n0;
      // row -> pixel
      int xl, yl, il;
      if (p.flat) {
        xl = r;
        yl = 0;
        il = 0;
      } else {
        xl = r % p.bw;
        yl = (r / p.bw) % p.bh;
        il = r / (p.bw * p.bh);
      }
      const int x = x0 + xl, y = y0 + yl, img_c = i0 + il;
      const bool row_ok = (x < d.x_end) && (y < p.H) && (img_c < p.NI);
      const int img = p.flat ? (p.rows_per_img > 0 ? x / p.rows_per_img : 0) : img_c;

      const bool ln_in = p.ln_rowstat != nullptr;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (ln_in && row_ok) {  // LayerNorm statistics of this thread's INPUT row (accumulated by the producing GEMM)
        const long long row = (static_cast<long long>(img_c) * p.H + y) * p.W + x;
        const float2 sq = *reinterpret_cast<const float2*>(p.ln_rowstat + 2 * row);
        ln_mean = sq.x * p.ln_inv_k;
        ln_rstd = rsqrtf(fmaxf(sq.y * p.ln_inv_k - ln_mean * ln_mean, 0.f) + p.ln_eps);
      }
      // (A) the previous unit's TMA stores have finished reading the panels (thread 64 waited before this barrier)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      auto issue_residual = [&]() {
        if (has_res && threadIdx.x == 64) {
          mbar_expect_tx(&res_bar, static_cast<uint32_t>(kBlockM * n_tile_out * 2));
          for (int pn = 0; pn < full_panels; ++pn)
            tma_load_4d(panels + pn * 16384, &tmR, &res_bar, oc0 + pn * 64, x0, y0, i0);
          if (rem) tma_load_4d(panels + full_panels * 16384, &tmRp, &res_bar, oc0 + full_panels * 64, x0, y0, i0);
        }
      };
      // the residual tile rides in the output panels; without split-K it is fetched while the MMAs still run
      if (p.splits == 1) issue_residual();
      if (gn_panel)
        for (int i = et; i < 4 * kGnSlots * 2; i += 256) gstat_s[i] = 0.f;
      // per-column epilogue vector: vec[quarter][col] = bias[col] + rowvec[image of the quarter's rows][col]
      for (int col = et; col < kPairN; col += 256) {
        const bool col_ok = n0 + col < p.N;
        const float bv = (col_ok && p.bias) ? p.bias[b_noff + n0 + col] : 0.f;
        if (ln_in) {  // slot 0: folded bias, slot 1: column sums of the gamma-scaled weights
          vec_s[col] = bv;
          vec_s[kPairN + col] = col_ok ? p.ln_colsum[b_noff + n0 + col] : 0.f;
          continue;
        }
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
          float rv = 0.f;
          if (col_ok && p.rowvec) {
            const int r4 = w4 * 32;
            int ximg;
            bool ok4;
            if (p.flat) {
              ximg = (x0 + r4) / p.rows_per_img;
              ok4 = x0 + r4 < d.x_end;
            } else {
              ximg = i0 + r4 / (p.bw * p.bh);
              ok4 = ximg < p.NI;
            }
            if (ok4) rv = p.rowvec[static_cast<long long>(ximg) * p.rowvec_ld + n0 + col];
          }
          vec_s[w4 * kPairN + col] = bv + rv;
        }
      }
      mbar_wait(&tmem_full_bar, ui & 1);
      tc_fence_after();
      if (ui == 0 && threadIdx.x == 64) PAIR_TRACE(4);
      // (B) vec_s visible to every epilogue warp
      asm volatile("bar.sync 1, 256;" ::: "memory");

      // ---- split-K: publish this CTA's partial tile; only the last arriver continues to the epilogue ----
      const int tile_id = d.tn * (2 * m_pairs) + d.tile_m;
      const long long tiles = static_cast<long long>(2 * m_pairs) * n_tiles;
      bool from_ws = false;
      bool skip = false;
      if (p.splits > 1) {
        float4* wp = reinterpret_cast<float4*>(p.ws_partial) +
                     (static_cast<long long>(d.z) * tiles + tile_id) * (kBlockM * kPairN / 4) + r;
#pragma unroll 1
        for (int c = 0; c < kPairNH; c += 16) {
          uint32_t v[16];
          tmem_ld_x16(t_row + c, v);
          tmem_ld_wait();
          const int cg = half * kPairNH + c;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 f = has_work ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            __stcg(wp + static_cast<long long>(cg / 4 + j) * kBlockM, f);
          }
        }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
          const int old = atomicAdd(p.ws_counter + tile_id, 1);
          splitk_last = (old == p.splits - 1) ? 1 : 0;
          if (splitk_last) p.ws_counter[tile_id] = 0;  // self-reset for the next launch
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!splitk_last) skip = true;
        else {
          __threadfence();
          from_ws = true;
          issue_residual();
        }
      }
      if (from_ws || skip) {
        // the accumulator is not needed any more: hand TMEM back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_empty);
      }
      if (!skip) {
        const float* vrow = (ln_in ? vec_s : vec_s + q * kPairN) + half * kPairNH;
        const float* srow = vec_s + kPairN + half * kPairNH;  // LayerNorm-folded GEMM: column sums
        float rs_acc = 0.f, rq_acc = 0.f;                     // producer side: (sum, sumsq) of this thread's output row
        auto vec16 = [&](int col, float (&b)[16]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 f = *reinterpret_cast<const float4*>(vrow + col + 4 * j);
            b[4 * j] = f.x; b[4 * j + 1] = f.y; b[4 * j + 2] = f.z; b[4 * j + 3] = f.w;
          }
        };
        auto pre16 = [&](int col, float (&a)[16]) {
          float b[16];
          vec16(col, b);
          if (ln_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 f = *reinterpret_cast<const float4*>(srow + col + 4 * j);
              a[4 * j] = (a[4 * j] - ln_mean * f.x) * ln_rstd + b[4 * j];
              a[4 * j + 1] = (a[4 * j + 1] - ln_mean * f.y) * ln_rstd + b[4 * j + 1];
              a[4 * j + 2] = (a[4 * j + 2] - ln_mean * f.z) * ln_rstd + b[4 * j + 2];
              a[4 * j + 3] = (a[4 * j + 3] - ln_mean * f.w) * ln_rstd + b[4 * j + 3];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] += b[j];
          }
        };
        // partial-tile sum for accumulator columns [cg, cg + 16) of this thread's row (split-K last arriver)
        auto load_ws = [&](int cg, float (&o)[16]) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = 0.f;
          for (int sp = 0; sp < p.splits; ++sp) {
            const float4* wp = reinterpret_cast<const float4*>(p.ws_partial) +
                               (static_cast<long long>(sp) * tiles + tile_id) * (kBlockM * kPairN / 4) +
                               static_cast<long long>(cg / 4) * kBlockM + r;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 f = __ldcg(wp + j * kBlockM);
              o[4 * j] += f.x; o[4 * j + 1] += f.y; o[4 * j + 2] += f.z; o[4 * j + 3] += f.w;
            }
          }
        };
        auto gn_accumulate = [&](const float (&o)[16], int ct) {  // ct: column inside the 320-wide tile
          const int col0 = n0 + ct;
          int nvalid = p.N - col0;
          if (nvalid > 16) nvalid = 16;
          if (nvalid <= 0) return;
          const int g_first = col0 / p.gn_cpg, g_last = (col0 + nvalid - 1) / p.gn_cpg;
          for (int g = g_first; g <= g_last; ++g) {
            const int lo = max(col0, g * p.gn_cpg) - col0, hi = min(col0 + nvalid, (g + 1) * p.gn_cpg) - col0;
            float sv = 0.f, qv = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j >= lo && j < hi && row_ok) {
                sv += o[j];
                qv += o[j] * o[j];
              }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              sv += __shfl_xor_sync(0xffffffffu, sv, off);
              qv += __shfl_xor_sync(0xffffffffu, qv, off);
            }
            const int img0 = __shfl_sync(0xffffffffu, img, 0);
            const int ok0 = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, 0);
            if (lane == 0 && (ok0 || sv != 0.f || qv != 0.f)) {
              float* w = p.gn_ws + (static_cast<long long>(img0) * p.gn_groups + g) * 2;
              atomicAdd(w, sv);
              atomicAdd(w + 1, qv);
            }
          }
        };
        if (has_res) {
          mbar_wait(&res_bar, res_it & 1);
          ++res_it;
        }
        // co: output column inside the tile's output panel row (0..n_tile_out), 16 wide
        auto finish_chunk = [&](int co, float (&o)[16]) {
          const int pn = co >> 6;
          uint8_t* pbase = panels + pn * 16384;
          uint4* d0;
          uint4* d1;
          if (pn < full_panels) {
            const int ch = (co & 63) >> 3;
            d0 = reinterpret_cast<uint4*>(pbase + r * 128 + ((ch ^ (r & 7)) << 4));
            d1 = reinterpret_cast<uint4*>(pbase + r * 128 + (((ch + 1) ^ (r & 7)) << 4));
          } else {
            d0 = reinterpret_cast<uint4*>(pbase + r * (rem * 2) + (co & 63) * 2);
            d1 = d0 + 1;
          }
          if (has_res) {
            const uint4 r0 = *d0, r1 = *d1;
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 f = Cvt<T>::unpack2(rr[j]);
              o[2 * j] += f.x;
              o[2 * j + 1] += f.y;
            }
          }
          if (p.gn_ws && !geglu && !gn_panel) gn_accumulate(o, co);
          if (p.rowstat_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              rs_acc += o[j];
              rq_acc += o[j] * o[j];
            }
          }
          uint4 w0, w1;
          w0.x = Cvt<T>::pack2(o[0], o[1]); w0.y = Cvt<T>::pack2(o[2], o[3]);
          w0.z = Cvt<T>::pack2(o[4], o[5]); w0.w = Cvt<T>::pack2(o[6], o[7]);
          w1.x = Cvt<T>::pack2(o[8], o[9]); w1.y = Cvt<T>::pack2(o[10], o[11]);
          w1.z = Cvt<T>::pack2(o[12], o[13]); w1.w = Cvt<T>::pack2(o[14], o[15]);
          *d0 = w0;
          *d1 = w1;
        };
        constexpr int GH = kPairNH / 2;  // GEGLU: value columns [0, 80), gate columns [80, 160) of this warp's half
        if (from_ws) {
          const int nloc = geglu ? GH : kPairNH;
#pragma unroll 1
          for (int c = 0; c < nloc; c += 16) {
            float o[16];
            if (geglu) {
              float a[16], g[16];
              load_ws(half * kPairNH + c, a);
              load_ws(half * kPairNH + GH + c, g);
              pre16(c, a);
              pre16(GH + c, g);
              epi_geglu16(o, a, g, p.alpha);  // stage by stage across the eight pairs (epilogue.cuh)
              finish_chunk(half * GH + c, o);
            } else {
              load_ws(half * kPairNH + c, o);
              pre16(c, o);
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] *= p.alpha;
              finish_chunk(half * kPairNH + c, o);
            }
          }
        } else if (geglu) {
          uint32_t va[2][16], vg[2][16];
          tmem_ld_x16(t_row, va[0]);
          tmem_ld_x16(t_row + GH, vg[0]);
          auto geglu_chunk = [&](int c, const uint32_t (&a)[16], const uint32_t (&g)[16]) {
            float o[16], av[16], gv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              av[j] = __uint_as_float(a[j]);
              gv[j] = __uint_as_float(g[j]);
            }
            pre16(c, av);
            pre16(GH + c, gv);
            epi_geglu16(o, av, gv, p.alpha);  // stage by stage across the eight pairs (epilogue.cuh)
            finish_chunk(half * GH + c, o);
          };
#pragma unroll 1
          for (int c = 0; c < GH; c += 32) {
            tmem_ld_wait();
            if (c + 16 < GH) {
              tmem_ld_x16(t_row + c + 16, va[1]);
              tmem_ld_x16(t_row + GH + c + 16, vg[1]);
            }
            geglu_chunk(c, va[0], vg[0]);
            if (c + 16 < GH) {
              tmem_ld_wait();
              if (c + 32 < GH) {
                tmem_ld_x16(t_row + c + 32, va[0]);
                tmem_ld_x16(t_row + GH + c + 32, vg[0]);
              }
              geglu_chunk(c + 16, va[1], vg[1]);
            }
          }
        } else {
          uint32_t v[2][16];
          tmem_ld_x16(t_row, v[0]);
          auto plain_chunk = [&](int c, const uint32_t (&a)[16]) {
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(a[j]);
            pre16(c, o);
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] *= p.alpha;
            finish_chunk(half * kPairNH + c, o);
          };
#pragma unroll 1
          for (int c = 0; c < kPairNH; c += 32) {
            tmem_ld_wait();
            if (c + 16 < kPairNH) tmem_ld_x16(t_row + c + 16, v[1]);
            plain_chunk(c, v[0]);
            if (c + 16 < kPairNH) {
              tmem_ld_wait();
              if (c + 32 < kPairNH) tmem_ld_x16(t_row + c + 32, v[0]);
              plain_chunk(c + 16, v[1]);
            }
          }
        }
        if (p.rowstat_out && row_ok) {
          const long long row = (static_cast<long long>(img_c) * p.H + y) * p.W + x;
          atomicAdd(p.rowstat_out + 2 * row, rs_acc);
          atomicAdd(p.rowstat_out + 2 * row + 1, rq_acc);
        }
        if (!from_ws) {
          // every tcgen05.ld of this warp has completed: hand TMEM back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(lead_empty);
        }
        fence_proxy_async_smem();
      }
      // (C) panels complete
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (!skip && threadIdx.x == 64) {
        for (int pn = 0; pn < full_panels; ++pn) tma_store_4d(&tmO, panels + pn * 16384, oc0 + pn * 64, x0, y0, i0);
        if (rem) tma_store_4d(&tmOp, panels + full_panels * 16384, oc0 + full_panels * 64, x0, y0, i0);
        tma_store_commit();
      }
      if (gn_panel) {
        // GroupNorm statistics of the finished tile from the smem panels: one 8-column chunk per lane over the
        // warp's 32 rows, reduced in smem, one global atomic per (image, group, statistic)
        const int img_t0 = p.flat ? (p.rows_per_img > 0 ? x0 / p.rows_per_img : 0) : i0;
        const int g_t0 = (p.gn_col0 + n0) / p.gn_cpg;
        if (!skip) {
          const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
          const int img_w = __shfl_sync(0xffffffffu, img, 0);
          const int cc = half * (kPairNH / 8) + lane;
          const int col0 = n0 + cc * 8;
          if (lane < kPairNH / 8 && okmask != 0 && col0 < p.N && img_w - img_t0 < 4) {
            float sv[8], qv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j] = qv[j] = 0.f;
            const uint8_t* pb = panels + (cc >> 3) * 16384;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              if ((okmask >> rr) & 1u) {
                const int row = q * 32 + rr;
                const uint4 u4 = *reinterpret_cast<const uint4*>(pb + row * 128 + (((cc & 7) ^ (row & 7)) << 4));
                const uint32_t uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = Cvt<T>::unpack2(uu[j]);
                  sv[2 * j] += f.x; qv[2 * j] += f.x * f.x;
                  sv[2 * j + 1] += f.y; qv[2 * j + 1] += f.y * f.y;
                }
              }
            }
            const int nval = min(8, p.N - col0);
            const int g0 = (p.gn_col0 + col0) / p.gn_cpg;
            const int bnd = (g0 + 1) * p.gn_cpg - (p.gn_col0 + col0);
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < nval) {
                if (j < bnd) { s0 += sv[j]; q0 += qv[j]; }
                else { s1 += sv[j]; q1 += qv[j]; }
              }
            float* gs = gstat_s + ((img_w - img_t0) * kGnSlots + (g0 - g_t0)) * 2;
            atomicAdd(gs, s0);
            atomicAdd(gs + 1, q0);
            if (bnd < nval) {
              atomicAdd(gs + 2, s1);
              atomicAdd(gs + 3, q1);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!skip)
          for (int i = et; i < 4 * kGnSlots * 2; i += 256) {
            const float v = gstat_s[i];
            if (v != 0.f) {
              const int il4 = i / (kGnSlots * 2), gl = (i >> 1) % kGnSlots;
              atomicAdd(p.gn_ws + (static_cast<long long>(img_t0 + il4) * p.gn_groups + g_t0 + gl) * 2 + (i & 1), v);
            }
          }
      }
      if (!skip && threadIdx.x == 64) {
        tma_store_wait_read0();
        if (ui == 0) PAIR_TRACE(5);
      }
    }
  }
  if (threadIdx.x == 64) PAIR_TRACE(6);

  // ---- teardown: both CTAs are done with TMEM and with each other's shared memory ---------------------
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kPairTmemCols);
  }
}

// Host side.  Returns 1 when the pair kernel cannot run this problem (caller falls back to the single-CTA kernel).
static bool gemm_pair_eligible(const EsGemm* g, const GemmKParams& kp, int m_tiles) {
  if (g->n % kPairN != 0 || g->out_fp32 || g->b_blocked || g->act == ES_ACT_SILU) return false;
  if (g->act == ES_ACT_GEGLU && (g->n / 2) % 8 != 0) return false;
  const int n_out = g->act == ES_ACT_GEGLU ? g->n / 2 : g->n;
  const bool aligned = (reinterpret_cast<uintptr_t>(g->out) & 15) == 0 && g->ldc % 8 == 0 && n_out % 8 == 0 &&
                       (!g->residual || ((reinterpret_cast<uintptr_t>(g->residual) & 15) == 0 && g->ldr % 8 == 0));
  if (!aligned) return false;
  if (kp.flat) {
    // a pair must not straddle two row segments (different weights) and segment tails are not clipped by the TMA store
    for (int s = 0; s < kp.nseg; ++s) {
      const int rows = kp.seg_row_start[s + 1] - kp.seg_row_start[s];
      if (kp.nseg > 1 && rows % (2 * kBlockM) != 0) return false;
    }
    if (g->rowvec && !(g->rows_per_img > 0 && g->rows_per_img % 32 == 0)) return false;
  } else {
    if (g->rowvec && (kp.bw * kp.bh) % 32 != 0) return false;
    // the two CTAs of a pair share one weight tile: both of their M tiles must lie in the same image segment
    const int tiles_per_img_group = kp.tiles_x * kp.tiles_y;
    for (int s = 0; s < kp.nseg && kp.nseg > 1; ++s) {
      const int t0 = kp.seg_row_start[s] / kp.bn * tiles_per_img_group;
      const int t1 = ceil_div(kp.seg_row_start[s + 1], kp.bn) * tiles_per_img_group;
      if (t1 > t0 && (t0 % 2 != 0 || (t1 % 2 != 0 && s != kp.nseg - 1))) return false;
    }
  }
  (void)m_tiles;
  return true;
}

template <typename T>
static int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmA2, GemmKParams& kp, int m_tiles,
                            const EsGemm* g, cudaStream_t stream) {
  // weight maps with 80-row boxes (each CTA loads its half of a 160-row UMMA operand)
  CUtensorMap tmB, tmB2;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(g->c1), static_cast<uint64_t>(g->taps), static_cast<uint64_t>(g->n_total_b)};
    uint64_t strides[3] = {0, static_cast<uint64_t>(g->c1) * 2, static_cast<uint64_t>(g->c1) * 2 * g->taps};
    uint32_t box[3] = {static_cast<uint32_t>(kBlockK), 1u, static_cast<uint32_t>(kPairNH / 2)};
    if (encode_tmap_16b(&tmB, g->b, 3, dims, strides, box)) return -3;
  }
  tmB2 = tmB;
  if (g->a2) {
    uint64_t dimsb[3] = {static_cast<uint64_t>(g->c2), 1, static_cast<uint64_t>(g->n_total_b2)};
    uint64_t stridesb[3] = {0, static_cast<uint64_t>(g->c2) * 2, static_cast<uint64_t>(g->c2) * 2};
    uint32_t boxb[3] = {static_cast<uint32_t>(kBlockK), 1u, static_cast<uint32_t>(kPairNH / 2)};
    if (encode_tmap_16b(&tmB2, g->b2, 3, dimsb, stridesb, boxb)) return -3;
  }
  CUtensorMap tmO = tmA, tmOp = tmA, tmR = tmA, tmRp = tmA;
  kp.n_out = g->act == ES_ACT_GEGLU ? g->n / 2 : g->n;
  kp.tma_epi = 1;
  {
    const int n_tile_out = g->act == ES_ACT_GEGLU ? kPairNH : kPairN;
    const int rem = n_tile_out % 64;
    auto make = [&](CUtensorMap* full, CUtensorMap* part, const void* base, long long ld) -> int {
      const uint64_t pitch = static_cast<uint64_t>(ld) * 2;
      uint64_t dims[4] = {static_cast<uint64_t>(kp.n_out), static_cast<uint64_t>(kp.W), static_cast<uint64_t>(kp.H),
                          static_cast<uint64_t>(kp.NI)};
      uint64_t strides[4] = {0, pitch, pitch * kp.W, pitch * kp.W * kp.H};
      uint32_t box[4] = {64u, static_cast<uint32_t>(kp.bw), static_cast<uint32_t>(kp.bh), static_cast<uint32_t>(kp.bn)};
      if (encode_tmap_16b(full, base, 4, dims, strides, box, true)) return -1;
      if (rem) {
        box[0] = static_cast<uint32_t>(rem);
        if (encode_tmap_16b(part, base, 4, dims, strides, box, false)) return -1;
      }
      return 0;
    };
    if (make(&tmO, &tmOp, g->out, g->ldc)) return -3;
    if (g->residual && make(&tmR, &tmRp, g->residual, g->ldr)) return -3;
  }
  const int m_pairs = (m_tiles + 1) / 2;
  const int n_tiles = g->n / kPairN;
  const int kb_total = kp.taps * kp.kblocks1 + kp.kblocks2;
  const int tiles = 2 * m_pairs * n_tiles;
  const int pair_units = m_pairs * n_tiles;
  const int max_clusters = 74;  // 148 SMs, one CTA per SM
  int splits = 1;
  if (g->workspace && g->split_k != 1) {
    if (g->split_k > 1) {
      splits = g->split_k;
    } else if (pair_units <= max_clusters / 2 && kb_total >= 16) {
      splits = max_clusters / pair_units;
      if (splits > kb_total / 6) splits = kb_total / 6;
      if (splits > 32) splits = 32;
      if (splits < 1) splits = 1;
    }
    const long long need = 65536 + static_cast<long long>(splits) * tiles * kBlockM * kPairN * 4;
    if (splits > 1 && (need > g->workspace_bytes || tiles > 16384)) splits = 1;
  }
  kp.splits = splits;
  kp.ws_counter = reinterpret_cast<int*>(g->workspace);
  kp.ws_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(g->workspace) + 65536);
  kp.stages = kPairStages;
  const int n_units = pair_units * splits;
  const int clusters = n_units < max_clusters ? n_units : max_clusters;
  auto kern = gemm_pair_kernel<T>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
    attr_done = true;
  }
  ES_CUDA(launch_kernel(kern, dim3(2 * clusters), dim3(kPairThreads), static_cast<size_t>(kPairSmem), stream, tmA, tmB,
                        tmA2, tmB2, tmO, tmOp, tmR, tmRp, kp, n_units, m_pairs, n_tiles));
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es
