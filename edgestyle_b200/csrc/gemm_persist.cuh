// es_gemm, persistent variant: one CTA per SM loops over output tiles with TWO accumulators in TMEM, so the epilogue
// of tile i (TMEM drain, bias / residual / GEGLU / folded LayerNorm, smem panels, TMA store) runs under the MMAs of
// tile i+1.
//
// Why (measured on B200, tools/gemm_trace.py, profiles/r2e_gemm_trace.txt): a one-tile-per-CTA launch spends, per 128 x 160
// tile of a K = 320 layer, 3.7 k cycles in set-up (barrier init, TMEM allocation, descriptor fetch, cold instruction
// cache), 2.7 k in the mainloop and 5.4 k in the epilogue (3.2 k of it the TMEM drain by ONE warp per sub-partition:
// 320 cycles per 16-column chunk, a single warp's dependent-issue latency) -- 12 k cycles of which 2.7 k use the tensor
// pipe; two CTAs per SM only overlap them pairwise.  Here the set-up is paid once per SM, sixteen epilogue warps (four per
// TMEM lane quarter, each a quarter of the tile's columns) drain four times as fast, and nothing but the MMA warp's own issue
// stream separates the tiles of one SM.
//
//   warp 0     : TMA producer over a ring of kStages x (A 16 KB + B BLOCK_N x 128 B); the ring runs on across tiles, so
//                the first K blocks of the next tile are in flight while the current one finishes.
//   warp 1     : TMEM allocation (2 x BLOCK_N columns) and single-thread tcgen05.mma issue (128 x BLOCK_N x 16);
//                accumulator stage = unit & 1, handed over with tmem_full / tmem_empty mbarriers.
//   warps 2..17: epilogue into DEDICATED panel smem (the ring belongs to the next tile), TMA store.
//   work units : (m tile, n tile) with m fastest, unit += gridDim.x; grid = min(units, SMs).
// No split-K here (the launcher only picks this kernel for grids that fill the machine).
#pragma once

namespace es {

#ifdef ES_GEMM_TRACE
// CTA 0: slots 16 + 12 * unit + k for its first four units
#define PS_TRACE(ui, k)                                                                   \
  do {                                                                                    \
    if (blockIdx.x == 0 && (ui) < 4) g_gemm_trace[16 + 12 * (ui) + (k)] = clock64();       \
  } while (0)
#else
#define PS_TRACE(ui, k) do {} while (0)
#endif

constexpr int kPsEpiWarps = 8;  // two per TMEM lane quarter (12 or 16 warps need < 128 registers: the epilogue spills and loses, measured): the drain is bound by a warp's dependent-issue latency
constexpr int kPsEpiThreads = 32 * kPsEpiWarps;
constexpr int kPsThreads = 64 + kPsEpiThreads;

template <int BLOCK_N>
struct PsCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPanels = (BLOCK_N + 63) / 64;
  static constexpr int kPanelBytes = kPanels * 16384;
  static constexpr int kVecBytes = 2 * 4 * BLOCK_N * 4;  // two units' epilogue vectors (the next one is staged early)
  static constexpr int kGnBytes = 4 * kGnSlots * 2 * 4;
  static constexpr int kFixed = kPanelBytes + kVecBytes + kGnBytes + 2048;
  static constexpr int kStagesRaw = (227 * 1024 - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmem = kStages * kStageBytes + kFixed;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N <= 256 ? 256 : 512;
};

struct PsUnit {
  int x0, y0, i0, x_end, b_noff, b2_noff, n0, tn, kb_total;
};

__device__ __forceinline__ PsUnit ps_decode(const GemmKParams& p, int u, int m_tiles, int block_n) {
  PsUnit d;
  const int tile_m = u % m_tiles;
  d.tn = u / m_tiles;
  d.n0 = d.tn * block_n;
  if (p.flat) {
    int g = 0;
#pragma unroll
    for (int s = 1; s < ES_MAX_SEG; ++s)
      if (s < p.nseg && tile_m >= p.seg_tile_start[s]) g = s;
    d.x0 = p.seg_row_start[g] + (tile_m - p.seg_tile_start[g]) * kBlockM;
    d.x_end = p.seg_row_start[g + 1];
    d.y0 = 0;
    d.i0 = 0;
    d.b_noff = p.seg_b_noff[g];
    d.b2_noff = p.seg_b2_noff[g];
  } else {
    const int tx = tile_m % p.tiles_x;
    const int ty = (tile_m / p.tiles_x) % p.tiles_y;
    const int tnn = tile_m / (p.tiles_x * p.tiles_y);
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
  d.kb_total = kb1 + ((p.kblocks2 > 0 && d.b2_noff >= 0) ? p.kblocks2 : 0);
  return d;
}

template <typename T, int BLOCK_N>
__global__ void __launch_bounds__(kPsThreads, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmOp,
                    const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmRp,
                    const GemmKParams p, const int n_units, const int m_tiles) {
  using Cfg = PsCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed on the SHARED-window address: going through uintptr_t would make
  // every later access a generic LD / ST (64-bit address arithmetic, no LDS / STS) -- measured in the epilogues
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* panels = smem + kStages * Cfg::kStageBytes;
  float* vec_s0 = reinterpret_cast<float*>(panels + Cfg::kPanelBytes);  // [2 units][4 lane quarters][BLOCK_N]
  float* gstat_s = vec_s0 + 2 * 4 * BLOCK_N;
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];   // MMA -> epilogue
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];  // 8 epilogue warps -> MMA
  __shared__ __align__(8) uint64_t res_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb1 = p.taps * p.kblocks1;
  pdl_launch_dependents();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kblocks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kPsEpiWarps);
    }
    mbar_init(&res_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // PDL: global memory written by the previous kernel may be touched from here on

  if (warp == 0) {
    // =============================== TMA producer ===========================================
    if (elect_one()) {
      int it = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const PsUnit d = ps_decode(p, u, m_tiles, BLOCK_N);
        for (int kb = 0; kb < d.kb_total; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          if (kb < kb1) {
            const int tap = kb / p.kblocks1;
            const int cb = kb - tap * p.kblocks1;
            int dx = 0, dy = 0;
            if (p.taps == 9) {
              dy = tap / 3 - 1;
              dx = tap % 3 - 1;
            }
            tma_load_3d(sb, &tmB, &full_bar[s], cb * kBlockK, tap, d.b_noff + d.n0);
            tma_load_4d(sa, &tmA, &full_bar[s], cb * kBlockK, d.x0 + dx, d.y0 + dy, d.i0);
          } else {
            const int cb = kb - kb1;
            tma_load_3d(sb, &tmB2, &full_bar[s], cb * kBlockK, 0, d.b2_noff + d.n0);
            tma_load_4d(sa, &tmA2, &full_bar[s], cb * kBlockK, d.x0, d.y0, d.i0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =============================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N, Cvt<T>::kFmt, 0, 0);
      int it = 0, ui = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
        const PsUnit d = ps_decode(p, u, m_tiles, BLOCK_N);
        const int as = ui & 1;
        mbar_wait(&tmem_empty_bar[as], ((ui >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        PS_TRACE(ui, 0);
        const uint32_t tacc = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < d.kb_total; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == 0) PS_TRACE(ui, 1);
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = smem_desc_sw128(sa + Cfg::kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[s]);
        }
        PS_TRACE(ui, 2);
        if (d.kb_total > 0) umma_commit(&tmem_full_bar[as]);
        else mbar_arrive(&tmem_full_bar[as]);
      }
    }
  } else {
    // =============================== epilogue (8 warps) =====================================
    const int q = warp & 3;             // TMEM lane quarter of this warp
    const int part = (warp - 2) >> 2;   // which part of the tile's column chunks (kPsEpiWarps / 4 parts)
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;    // 0..kPsEpiThreads-1
    const bool geglu = p.act == ES_ACT_GEGLU;
    const bool has_res = p.residual != nullptr;
    constexpr int GH = BLOCK_N / 2;     // GEGLU: value columns [0, GH), gate columns [GH, BLOCK_N)
    const int n_tile_out = geglu ? GH : BLOCK_N;
    const int full_panels = n_tile_out / 64;
    const int rem = n_tile_out % 64;
    const int n_chunks = n_tile_out / 16;  // 16-column output chunks of a tile, split between the two warp halves
    constexpr int kParts = kPsEpiWarps / 4;
    const int c_begin = (n_chunks * part) / kParts;
    const int c_end = (n_chunks * (part + 1)) / kParts;
    const bool gn_panel = p.gn_ws != nullptr;
    const bool ln_in = p.ln_rowstat != nullptr;
    int ui = 0, res_it = 0;
    // Everything a unit's epilogue needs from global memory besides the accumulator -- the LayerNorm statistics of the
    // thread's input row and the per-column vector (bias / folded bias + column sums / bias + time-embedding row) -- is
    // FETCHED ONE UNIT AHEAD (folded-LayerNorm GEMMs: before the previous unit's drain; the others: at its end, under the
    // TMA store's read of the panels), the vector into the other half of vec_s: two dependent global loads (~650 cycles each) sat in front of every unit's drain
    // (tools/persist_trace.py), and an epilogue warp has nothing else to issue meanwhile.  (Holding the prefetched
    // values in registers across the drain instead spills: 131 vs 110 us on the GEGLU GEMM.)
    constexpr int kVPT = (BLOCK_N + kPsEpiThreads - 1) / kPsEpiThreads;  // columns of the vector staged by one thread
    struct UnitRow {
      int x0, y0, i0, n0, oc0, b_noff, x_end, img;
      long long row;
      bool row_ok;
    };
    auto unit_row = [&](int u) {
      const PsUnit d = ps_decode(p, u, m_tiles, BLOCK_N);
      UnitRow ur;
      ur.x0 = d.x0, ur.y0 = d.y0, ur.i0 = d.i0, ur.n0 = d.n0, ur.b_noff = d.b_noff, ur.x_end = d.x_end;
      ur.oc0 = geglu ? d.tn * GH : d.n0;
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
      const int x = d.x0 + xl, y = d.y0 + yl, img_c = d.i0 + il;
      ur.row_ok = (x < d.x_end) && (y < p.H) && (img_c < p.NI);
      ur.img = p.flat ? (p.rows_per_img > 0 ? x / p.rows_per_img : 0) : img_c;
      ur.row = (static_cast<long long>(img_c) * p.H + y) * p.W + x;
      return ur;
    };
    auto fetch_ln = [&](const UnitRow& ur) {  // (sum, sumsq) of this thread's INPUT row (accumulated by the producing GEMM)
      return (ln_in && ur.row_ok) ? *reinterpret_cast<const float2*>(p.ln_rowstat + 2 * ur.row) : make_float2(0.f, 0.f);
    };
    // per-column epilogue vector: vec[quarter][col] = bias[col] + rowvec[image of the quarter's rows][col]; with a folded
    // LayerNorm slot 0 = folded bias, slot 1 = column sums of the gamma-scaled weights
    auto fetch_vec = [&](const UnitRow& ur, float (&v)[kVPT][4]) {
#pragma unroll
      for (int k = 0; k < kVPT; ++k) {
        const int col = et + k * kPsEpiThreads;
        v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.f;
        if (col >= BLOCK_N) continue;
        const bool col_ok = ur.n0 + col < p.N;
        const float bv = (col_ok && p.bias) ? p.bias[ur.b_noff + ur.n0 + col] : 0.f;
        if (ln_in) {
          v[k][0] = bv;
          v[k][1] = col_ok ? p.ln_colsum[ur.b_noff + ur.n0 + col] : 0.f;
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
              ximg = (ur.x0 + r4) / p.rows_per_img;
              ok4 = ur.x0 + r4 < ur.x_end;
            } else {
              ximg = ur.i0 + r4 / (p.bw * p.bh);
              ok4 = ximg < p.NI;
            }
            if (ok4) rv = p.rowvec[static_cast<long long>(ximg) * p.rowvec_ld + ur.n0 + col];
          }
          v[k][w4] = bv + rv;
        }
      }
    };
    auto park_vec = [&](const float (&v)[kVPT][4], float* dst) {
#pragma unroll
      for (int k = 0; k < kVPT; ++k) {
        const int col = et + k * kPsEpiThreads;
        if (col >= BLOCK_N) continue;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4)
          if (!ln_in || w4 < 2) dst[w4 * BLOCK_N + col] = v[k][w4];
      }
    };
    UnitRow cur = unit_row(blockIdx.x < n_units ? blockIdx.x : 0);
    float2 ln_cur = make_float2(0.f, 0.f);
    if (blockIdx.x < n_units) {
      float v0[kVPT][4];
      ln_cur = fetch_ln(cur);
      fetch_vec(cur, v0);
      park_vec(v0, vec_s0);
    }
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      const int as = ui & 1;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      float* vec_s = vec_s0 + (ui & 1) * 4 * BLOCK_N;
      const int x0 = cur.x0, y0 = cur.y0, i0 = cur.i0, n0 = cur.n0, oc0 = cur.oc0;
      const bool row_ok = cur.row_ok;
      const int img = cur.img;
      const long long row = cur.row;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (ln_in && row_ok) {
        ln_mean = ln_cur.x * p.ln_inv_k;
        ln_rstd = rsqrtf(fmaxf(ln_cur.y * p.ln_inv_k - ln_mean * ln_mean, 0.f) + p.ln_eps);
      }
      // Folded-LayerNorm GEMMs (every transformer projection: short K, the epilogue IS the unit): the next unit's row
      // statistics (two registers) and vector (cp.async straight into the other half of vec_s, no registers) are
      // requested before this unit's drain
      const bool have_next = u + static_cast<int>(gridDim.x) < n_units;
      float2 ln_next = make_float2(0.f, 0.f);
      if (have_next && ln_in) {
        const UnitRow nx = unit_row(u + gridDim.x);
        ln_next = fetch_ln(nx);
        float* dst = vec_s0 + ((ui + 1) & 1) * 4 * BLOCK_N;
#pragma unroll
        for (int k = 0; k < kVPT; ++k) {
          const int col = et + k * kPsEpiThreads;
          if (col < BLOCK_N) {
            const bool col_ok = nx.n0 + col < p.N;
            const int gc = col_ok ? nx.b_noff + nx.n0 + col : 0;
            cp_async_4(dst + col, p.bias ? p.bias + gc : p.ln_colsum + gc, col_ok && p.bias != nullptr);
            cp_async_4(dst + BLOCK_N + col, p.ln_colsum + gc, col_ok);
          }
        }
        cp_async_commit();
      }
      // (A) the previous unit's TMA stores have finished reading the panels (thread 64 waited before this barrier)
      asm volatile("bar.sync 1, %0;" ::"n"(kPsEpiThreads) : "memory");
      if (threadIdx.x == 64) PS_TRACE(ui, 3);
      if (has_res && warp == 2 && elect_one()) {  // the residual tile rides in the output panels, fetched while the MMAs run
        mbar_expect_tx(&res_bar, static_cast<uint32_t>(kBlockM * n_tile_out * 2));
        for (int pn = 0; pn < full_panels; ++pn) tma_load_4d(panels + pn * 16384, &tmR, &res_bar, oc0 + pn * 64, x0, y0, i0);
        if (rem) tma_load_4d(panels + full_panels * 16384, &tmRp, &res_bar, oc0 + full_panels * 64, x0, y0, i0);
      }
      if (gn_panel)
        for (int i = et; i < 4 * kGnSlots * 2; i += kPsEpiThreads) gstat_s[i] = 0.f;
      if (threadIdx.x == 64) PS_TRACE(ui, 4);
      mbar_wait(&tmem_full_bar[as], (ui >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64) PS_TRACE(ui, 5);
      // (B) vec_s visible to every epilogue warp
      asm volatile("bar.sync 1, %0;" ::"n"(kPsEpiThreads) : "memory");
      if (has_res) {
        mbar_wait(&res_bar, res_it & 1);
        ++res_it;
      }
      const float* vrow = ln_in ? vec_s : vec_s + q * BLOCK_N;
      const float* srow = vec_s + BLOCK_N;  // LayerNorm-folded GEMM: column sums
      uint64_t rs2 = pk2(0.f, 0.f), rq2 = pk2(0.f, 0.f);  // producer side: (sum, sumsq) of this thread's output row
      // accumulator columns [col, col + 16) -> pre-activation values (packed fp32 pairs: epilogue.cuh)
      auto pre16 = [&](int col, float (&a)[16]) {
        if (ln_in) epi_ln_vec16(a, vrow + col, srow + col, ln_mean, ln_rstd);
        else epi_add_vec16(a, vrow + col);
      };
      // co: output column inside the tile (0..n_tile_out), 16 wide -> this thread's 32 bytes of the smem panel
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
          epi_add_res16<T>(o, rr);
        }
        if (p.rowstat_out) epi_rowstat16(o, rs2, rq2);
        uint4 w0, w1;
        w0.x = Cvt<T>::pack2(o[0], o[1]); w0.y = Cvt<T>::pack2(o[2], o[3]);
        w0.z = Cvt<T>::pack2(o[4], o[5]); w0.w = Cvt<T>::pack2(o[6], o[7]);
        w1.x = Cvt<T>::pack2(o[8], o[9]); w1.y = Cvt<T>::pack2(o[10], o[11]);
        w1.z = Cvt<T>::pack2(o[12], o[13]); w1.w = Cvt<T>::pack2(o[14], o[15]);
        *d0 = w0;
        *d1 = w1;
      };
      if (geglu) {
        auto ld2 = [&](int ci, uint32_t (&a)[16], uint32_t (&g)[16]) {
          tmem_ld_x16(t_row + ci * 16, a);
          tmem_ld_x16(t_row + GH + ci * 16, g);
        };
        auto chunk = [&](int ci, const uint32_t (&a)[16], const uint32_t (&g)[16]) {
          float o[16], av[16], gv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            av[j] = __uint_as_float(a[j]);
            gv[j] = __uint_as_float(g[j]);
          }
          pre16(ci * 16, av);
          pre16(GH + ci * 16, gv);
          epi_geglu16(o, av, gv, p.alpha);
          finish_chunk(ci * 16, o);
        };
        uint32_t va[2][16], vg[2][16];
        if (c_begin < c_end) ld2(c_begin, va[0], vg[0]);
#pragma unroll 1
        for (int ci = c_begin; ci < c_end; ci += 2) {
          tmem_ld_wait();
          if (ci + 1 < c_end) ld2(ci + 1, va[1], vg[1]);
          chunk(ci, va[0], vg[0]);
          if (ci + 1 < c_end) {
            tmem_ld_wait();
            if (ci + 2 < c_end) ld2(ci + 2, va[0], vg[0]);
            chunk(ci + 1, va[1], vg[1]);
          }
        }
      } else {
        auto chunk = [&](int ci, const uint32_t (&a)[16]) {
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(a[j]);
          pre16(ci * 16, o);
          epi_scale16(o, p.alpha);
          finish_chunk(ci * 16, o);
        };
        uint32_t v[2][16];
        if (c_begin < c_end) tmem_ld_x16(t_row + c_begin * 16, v[0]);
#pragma unroll 1
        for (int ci = c_begin; ci < c_end; ci += 2) {
          tmem_ld_wait();
          if (ci + 1 < c_end) tmem_ld_x16(t_row + (ci + 1) * 16, v[1]);
          chunk(ci, v[0]);
          if (ci + 1 < c_end) {
            tmem_ld_wait();
            if (ci + 2 < c_end) tmem_ld_x16(t_row + (ci + 2) * 16, v[0]);
            chunk(ci + 1, v[1]);
          }
        }
      }
      // every tcgen05.ld of this warp has completed: hand the accumulator back to the MMA issuer
      if (threadIdx.x == 64) PS_TRACE(ui, 6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);

      if (p.rowstat_out && row_ok) {
        float rs_a, rs_b, rq_a, rq_b;
        upk2(rs2, rs_a, rs_b);
        upk2(rq2, rq_a, rq_b);
        atomicAdd(p.rowstat_out + 2 * row, rs_a + rs_b);
        atomicAdd(p.rowstat_out + 2 * row + 1, rq_a + rq_b);
      }
      fence_proxy_async_smem();
      // (C) panels complete
      asm volatile("bar.sync 1, %0;" ::"n"(kPsEpiThreads) : "memory");
      if (threadIdx.x == 64) PS_TRACE(ui, 7);
      if (warp == 2 && elect_one()) {  // (elect.sync, not a thread-index test, around TMA instructions)
        for (int pn = 0; pn < full_panels; ++pn) tma_store_4d(&tmO, panels + pn * 16384, oc0 + pn * 64, x0, y0, i0);
        if (rem) tma_store_4d(&tmOp, panels + full_panels * 16384, oc0 + full_panels * 64, x0, y0, i0);
        tma_store_commit();
      }
      if (gn_panel) {
        // GroupNorm statistics of the finished tile from the smem panels: one 8-column chunk per lane over the warp's 32
        // rows (the two warps of a lane quarter take alternate chunks), reduced in smem, one global atomic per
        // (image, group, statistic)
        const int img_t0 = p.flat ? (p.rows_per_img > 0 ? x0 / p.rows_per_img : 0) : i0;
        const int g_t0 = (p.gn_col0 + n0) / p.gn_cpg;
        const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
        const int img_w = __shfl_sync(0xffffffffu, img, 0);
        const int cc = part * 32 + lane;  // 8-column chunk of the tile handled by this lane
        const int col0 = n0 + cc * 8;
        if (cc < BLOCK_N / 8 && okmask != 0 && col0 < p.N && img_w - img_t0 < 4) {
          float sv[8], qv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) sv[j] = qv[j] = 0.f;
          const int pn = cc >> 3;
          const uint8_t* pb = panels + pn * 16384;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            if ((okmask >> rr) & 1u) {
              const int prow = q * 32 + rr;
              const uint4 u4 = (pn < full_panels)
                                   ? *reinterpret_cast<const uint4*>(pb + prow * 128 + (((cc & 7) ^ (prow & 7)) << 4))
                                   : *reinterpret_cast<const uint4*>(pb + prow * (rem * 2) + (cc & 7) * 16);
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
        asm volatile("bar.sync 1, %0;" ::"n"(kPsEpiThreads) : "memory");
        for (int i = et; i < 4 * kGnSlots * 2; i += kPsEpiThreads) {
          const float v = gstat_s[i];
          if (v != 0.f) {
            const int il4 = i / (kGnSlots * 2), gl = (i >> 1) % kGnSlots;
            atomicAdd(p.gn_ws + (static_cast<long long>(img_t0 + il4) * p.gn_groups + g_t0 + gl) * 2 + (i & 1), v);
          }
        }
      }
      if (threadIdx.x == 64) PS_TRACE(ui, 8);
      // the next unit's row statistics and vector: the loads run while the TMA store of this unit reads the panels; the
      // vector goes to the OTHER half of vec_s (slow warps may still read this unit's half; barrier (B) of the next unit
      // publishes it)
      if (have_next) {
        cur = unit_row(u + gridDim.x);
        if (ln_in) {
          ln_cur = ln_next;
          cp_async_wait_all();
        } else {
          float vn[kVPT][4];
          fetch_vec(cur, vn);
          park_vec(vn, vec_s0 + ((ui + 1) & 1) * 4 * BLOCK_N);
        }
      }
      if (warp == 2 && elect_one()) tma_store_wait_read0();  // (the same thread: elect.sync is deterministic per mask)
      if (threadIdx.x == 64) PS_TRACE(ui, 9);
    }
  }

  // ---- teardown --------------------------------------------------------------------------------
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Host side.  false when the persistent kernel cannot run this problem (the caller falls back to the one-tile kernels).
template <int BLOCK_N>
static bool gemm_persist_eligible(const EsGemm* g, const GemmKParams& kp) {
  if (g->out_fp32 || g->b_blocked || g->act == ES_ACT_SILU || g->split_k > 1) return false;
  if (g->act == ES_ACT_GEGLU && (g->n % BLOCK_N != 0 || (BLOCK_N / 2) % 16 != 0)) return false;
  const int n_out = g->act == ES_ACT_GEGLU ? g->n / 2 : g->n;
  const bool aligned = (reinterpret_cast<uintptr_t>(g->out) & 15) == 0 && g->ldc % 8 == 0 && n_out % 8 == 0 &&
                       (!g->residual || ((reinterpret_cast<uintptr_t>(g->residual) & 15) == 0 && g->ldr % 8 == 0));
  if (!aligned) return false;
  if (kp.flat) {
    // segment tails are not clipped by the TMA store: every segment but the last must end on a tile boundary
    for (int s = 0; s + 1 < kp.nseg; ++s)
      if ((kp.seg_row_start[s + 1] - kp.seg_row_start[s]) % kBlockM != 0) return false;
    if (g->rowvec && !(g->rows_per_img > 0 && g->rows_per_img % 32 == 0)) return false;
  } else {
    if (g->rowvec && (kp.bw * kp.bh) % 32 != 0) return false;
  }
  if (g->gn_ws) {
    const int cpg = g->gn_cpg > 0 ? g->gn_cpg : (g->gn_groups > 0 ? g->n / g->gn_groups : 0);
    if (g->act == ES_ACT_GEGLU || cpg < 8 || BLOCK_N / cpg + 2 > kGnSlots) return false;
  }
  return true;
}

template <typename T, int BLOCK_N>
static int launch_gemm_persist(const CUtensorMap& tmA, const CUtensorMap& tmA2, GemmKParams& kp, int m_tiles,
                               const EsGemm* g, cudaStream_t stream) {
  using Cfg = PsCfg<BLOCK_N>;
  CUtensorMap tmB, tmB2;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(g->c1), static_cast<uint64_t>(g->taps), static_cast<uint64_t>(g->n_total_b)};
    uint64_t strides[3] = {0, static_cast<uint64_t>(g->c1) * 2, static_cast<uint64_t>(g->c1) * 2 * g->taps};
    uint32_t box[3] = {static_cast<uint32_t>(kBlockK), 1u, static_cast<uint32_t>(BLOCK_N)};
    if (encode_tmap_16b(&tmB, g->b, 3, dims, strides, box)) return -3;
  }
  tmB2 = tmB;
  if (g->a2) {
    uint64_t dimsb[3] = {static_cast<uint64_t>(g->c2), 1, static_cast<uint64_t>(g->n_total_b2)};
    uint64_t stridesb[3] = {0, static_cast<uint64_t>(g->c2) * 2, static_cast<uint64_t>(g->c2) * 2};
    uint32_t boxb[3] = {static_cast<uint32_t>(kBlockK), 1u, static_cast<uint32_t>(BLOCK_N)};
    if (encode_tmap_16b(&tmB2, g->b2, 3, dimsb, stridesb, boxb)) return -3;
  }
  CUtensorMap tmO = tmA, tmOp = tmA, tmR = tmA, tmRp = tmA;
  kp.n_out = g->act == ES_ACT_GEGLU ? g->n / 2 : g->n;
  kp.tma_epi = 1;
  {
    const int n_tile_out = g->act == ES_ACT_GEGLU ? BLOCK_N / 2 : BLOCK_N;
    const int rem = n_tile_out % 64;
    auto make = [&](CUtensorMap* full, CUtensorMap* part, const void* base, long long ld) -> int {
      const uint64_t pitch = static_cast<uint64_t>(ld) * 2;
      uint64_t dims[4] = {static_cast<uint64_t>(kp.n_out), static_cast<uint64_t>(kp.W), static_cast<uint64_t>(kp.H),
                          static_cast<uint64_t>(kp.NI)};
      uint64_t strides[4] = {0, pitch, pitch * kp.W, pitch * kp.W * kp.H};
      uint32_t box[4] = {64u, static_cast<uint32_t>(kp.bw), static_cast<uint32_t>(kp.bh), static_cast<uint32_t>(kp.bn)};
      if (n_tile_out >= 64 && encode_tmap_16b(full, base, 4, dims, strides, box, true)) return -1;
      if (rem) {
        box[0] = static_cast<uint32_t>(rem);
        if (encode_tmap_16b(part, base, 4, dims, strides, box, false)) return -1;
      }
      return 0;
    };
    if (make(&tmO, &tmOp, g->out, g->ldc)) return -3;
    if (g->residual && make(&tmR, &tmRp, g->residual, g->ldr)) return -3;
  }
  if ((g->ln_rowstat || g->rowstat_out) && kp.flat && kp.nseg > 1)
    for (int sgi = 0; sgi + 1 < kp.nseg; ++sgi)
      ES_CHECK((kp.seg_row_start[sgi + 1] - kp.seg_row_start[sgi]) % kBlockM == 0,
               "es_gemm: folded LayerNorm needs row segments that are multiples of 128 rows");
  const int n_tiles = ceil_div(g->n, BLOCK_N);
  const int n_units = m_tiles * n_tiles;
  kp.splits = 1;
  kp.stages = Cfg::kStages;
  // ES_PERSIST_CTAS (read once): fewer CTAs than SMs leave room for the other streams' kernels next to a persistent GEMM
  static const int sms = [] {
    const char* e = getenv("ES_PERSIST_CTAS");
    const int v = e ? atoi(e) : 148;
    return v >= 1 && v <= 148 ? v : 148;
  }();
  const int ctas = n_units < sms ? n_units : sms;
  auto kern = gemm_persist_kernel<T, BLOCK_N>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    attr_done = true;
  }
  ES_CUDA(launch_kernel(kern, dim3(ctas), dim3(kPsThreads), static_cast<size_t>(Cfg::kSmem), stream, tmA, tmB, tmA2, tmB2,
                        tmO, tmOp, tmR, tmRp, kp, n_units, m_tiles));
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es
