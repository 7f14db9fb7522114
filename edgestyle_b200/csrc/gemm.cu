// es_gemm: warp-specialised tcgen05/TMEM implicit-GEMM (conv3x3 s1 / conv1x1 / Linear) for sm_100a.
//
//   warp 0      : TMA producer (one elected lane) -- A tile via a 4-D tensor map over the NHWC
//                 activation (box = 64 channels x bw x bh x bn pixels; the 9 taps of a 3x3 are
//                 shifted boxes, out-of-bounds rows are zero-filled by TMA == the conv padding),
//                 B tile via a 3-D map over [n][tap][c].
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x BLOCK_N x 16 per MMA).
//   warps 2..5  : epilogue -- tcgen05.ld the fp32 accumulator (one row per thread), fuse
//                 bias / per-image vector / GEGLU / alpha / residual, store 16 B vectors.
//   smem ring   : `stages` x (A 16 KB + B BLOCK_N*128 B), SWIZZLE_128B, full/empty mbarriers.
//
// Replaces the cuDNN/cuBLAS dispatches listed at include/edgestyle_b200.h (es_gemm).
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "epilogue.cuh"

namespace es {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kGemmThreads = 192;
constexpr int kGnSlots = 34;  // GroupNorm groups one output tile can touch (fused statistics), per image of the tile

// Optional event trace (clock64 stamps of CTA (1,0,0)) for pipeline analysis: build with -DES_GEMM_TRACE.
#ifdef ES_GEMM_TRACE
__device__ long long g_gemm_trace[64];
#define GEMM_TRACE(slot)                                                                        \
  do {                                                                                          \
    if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0) g_gemm_trace[slot] = clock64();  \
  } while (0)
#else
#define GEMM_TRACE(slot) do {} while (0)
#endif

struct GemmKParams {
  int W, H, NI;
  int bw, bh, bn;
  int tiles_x, tiles_y;
  int N;
  int taps;
  int kblocks1, kblocks2;
  int nseg;
  int seg_row_start[ES_MAX_SEG + 1];
  int seg_tile_start[ES_MAX_SEG + 1];
  int seg_b_noff[ES_MAX_SEG];
  int seg_b2_noff[ES_MAX_SEG];
  int flat;
  const float* bias;
  const float* rowvec;
  int rows_per_img;
  int rowvec_ld;
  const void* residual;
  long long ldr;
  int act;
  float alpha;
  void* out;
  long long ldc;
  int out_fp32;
  int stages;
  int splits;            // split-K factor (grid.z)
  int k_rotate;          // 1: every tile walks its K blocks from a different starting block (see produce_b)
  int coop;              // 1: cooperative split-K -- every split CTA reduces and finishes its own column slice of the tile
  float* ws_partial;     // [splits][tiles][128][BLOCK_N] fp32 partial accumulators
  int* ws_counter;       // [tiles] arrival counters (zero on entry, reset by the finishing CTA)
  float* gn_ws;          // optional: GroupNorm statistics of the OUTPUT accumulated here, [img][groups][2] (sum, sumsq)
  int gn_cpg, gn_groups; // channels per group, groups
  int gn_col0;           // channel of the normalised tensor that column 0 of this GEMM writes (concat halves)
  int b_blocked;         // weights stored K-block-major: [tap*kblocks + cb][n][64] (contiguous 128 B x BLOCK_N tiles)
  int tma_epi;           // 1: epilogue goes regs -> swizzled smem panels -> TMA store (residual via TMA load)
  int n_out;             // output columns in total (N, or N/2 for GEGLU)
  // LayerNorm folded around the GEMM (see EsGemm): producer side accumulates per-row (sum, sumsq) of the OUTPUT,
  // consumer side normalises with the statistics of its INPUT rows: out = rstd (acc - mean colsum[n]) + bias'[n]
  const void* prefetch;      // weights of the next GEMM of the stream: pulled into L2 while this kernel runs
  long long prefetch_bytes;
  float* rowstat_out;
  const float* ln_rowstat;
  const float* ln_colsum;
  float ln_inv_k, ln_eps;
};

// EXT selects the epilogue family at compile time: folded LayerNorm (row statistics of the output / normalisation by the
// statistics of the input rows) and the SiLU activation are separate instantiations, so the plain one keeps the lean
// epilogue (measured: a run-time SiLU select in the common epilogue costs 0.1 ms per step).
template <typename T, int BLOCK_N, int EXT>  // EXT: 0 lean epilogue, 1 folded LayerNorm, 2 SiLU activation
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
            const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmOp,
            const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmRp,
            const GemmKParams p) {
  constexpr bool LN = EXT == 1;
  constexpr bool SILU = EXT == 2;
  constexpr int kABytes = kBlockM * kBlockK * 2;
  constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  constexpr int kMaxStages = 12;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B, computed on the SHARED-window address: going through uintptr_t would make
  // every later access a generic LD / ST (64-bit address arithmetic, no LDS / STS) -- measured in the epilogues
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ __align__(8) uint64_t res_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int splitk_last;

  pdl_launch_dependents();
  if (threadIdx.x == 0) GEMM_TRACE(0);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;

  // ---- tile decode -------------------------------------------------------------------------
  const int tile_m = blockIdx.x;
  const int n0 = blockIdx.y * BLOCK_N;
  int x0, y0, i0, x_end, b_noff = 0, b2_noff = -1;
  if (p.flat) {
    int g = 0;
#pragma unroll
    for (int s = 1; s < ES_MAX_SEG; ++s)
      if (s < p.nseg && tile_m >= p.seg_tile_start[s]) g = s;
    x0 = p.seg_row_start[g] + (tile_m - p.seg_tile_start[g]) * kBlockM;
    x_end = p.seg_row_start[g + 1];
    y0 = 0;
    i0 = 0;
    b_noff = p.seg_b_noff[g];
    b2_noff = p.seg_b2_noff[g];
  } else {
    const int tx = tile_m % p.tiles_x;
    const int ty = (tile_m / p.tiles_x) % p.tiles_y;
    const int tn = tile_m / (p.tiles_x * p.tiles_y);
    x0 = tx * p.bw;
    y0 = ty * p.bh;
    i0 = tn * p.bn;
    x_end = p.W;
    // implicit conv: segments are IMAGE ranges (conv LoRA: one fused weight copy / one `up` matrix per image segment)
    int g = 0;
#pragma unroll
    for (int s = 1; s < ES_MAX_SEG; ++s)
      if (s < p.nseg && i0 >= p.seg_row_start[s]) g = s;
    b_noff = p.seg_b_noff[g];
    b2_noff = p.seg_b2_noff[g];
  }
  const int kb1 = p.taps * p.kblocks1;
  const int kb_total = kb1 + ((p.kblocks2 > 0 && b2_noff >= 0) ? p.kblocks2 : 0);
  // split-K: this CTA owns k-blocks [kb_begin, kb_end)
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  const int kb_begin = blockIdx.z * kb_per;
  const int kb_end = min(kb_total, kb_begin + kb_per);
  const bool has_work = kb_end > kb_begin;

  // ---- one-time setup ----------------------------------------------------------------------
  // TMA load of k-block kb into its ring slot (producer thread only)
  // TMA load of k-block kb into its ring slot (producer thread only).  The weight tile (B) does not depend on the
  // previous kernel of the stream, the activation tile (A) does: under programmatic dependent launch the first ring
  // fill issues the B loads before griddepcontrol.wait and the A loads after it.
  // Small-M launches: the few A tiles (and B tiles) of a K block are wanted by MANY CTAs at the same moment, and L2
  // serves the requests for one line one after the other (measured: 940 - 4600 cycles per K block on the 8x8-level
  // convolutions whatever the ring depth, tools/smallm_stages.py).  With k_rotate every (m, n) tile starts its K loop at
  // a different block and wraps around, so concurrent CTAs read different lines; the MMA warp just accumulates the
  // stages in ring order (the fp32 summation order differs per tile, deterministically).
  const int kb_len = max(kb_end - kb_begin, 1);
  const int k_rot = p.k_rotate ? static_cast<int>((blockIdx.x * 7u + blockIdx.y * 13u) % static_cast<unsigned>(kb_len)) : 0;
  auto rotated = [&](int kb) {
    int r = kb + k_rot;
    return r >= kb_end ? r - kb_len : r;
  };
  auto produce_b = [&](int kb_seq) {
    const int kb = rotated(kb_seq);
    const int it = kb_seq - kb_begin;
    const int s = it % stages;
    const uint32_t ph = (it / stages) & 1;
    mbar_wait(&empty_bar[s], ph ^ 1);
    uint8_t* sb = smem + s * kStageBytes + kABytes;
    mbar_expect_tx(&full_bar[s], kStageBytes);
    if (kb < kb1) {
      const int tap = kb / p.kblocks1;
      const int cb = kb - tap * p.kblocks1;
      if (p.b_blocked) tma_load_3d(sb, &tmB, &full_bar[s], 0, b_noff + n0, kb);
      else tma_load_3d(sb, &tmB, &full_bar[s], cb * kBlockK, tap, b_noff + n0);
    } else {
      tma_load_3d(sb, &tmB2, &full_bar[s], (kb - kb1) * kBlockK, 0, b2_noff + n0);
    }
  };
  auto produce_a = [&](int kb_seq) {
    const int kb = rotated(kb_seq);
    const int it = kb_seq - kb_begin;
    const int s = it % stages;
    uint8_t* sa = smem + s * kStageBytes;
    if (kb < kb1) {
      const int tap = kb / p.kblocks1;
      const int cb = kb - tap * p.kblocks1;
      int dx = 0, dy = 0;
      if (p.taps == 9) {
        dy = tap / 3 - 1;
        dx = tap % 3 - 1;
      }
      tma_load_4d(sa, &tmA, &full_bar[s], cb * kBlockK, x0 + dx, y0 + dy, i0);
    } else {
      tma_load_4d(sa, &tmA2, &full_bar[s], (kb - kb1) * kBlockK, x0, y0, i0);  // centre tap (1x1)
    }
  };
  auto produce = [&](int kb) {
    produce_b(kb);
    produce_a(kb);
  };
  const int kb_prefill = min(kb_end, kb_begin + stages);
  if (warp == 2 && lane == 0)  // constant data: no need to wait for the previous kernel
    l2_prefetch_slice(p.prefetch, p.prefetch_bytes, (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x,
                      gridDim.x * gridDim.y * gridDim.z);
  // (elect.sync, not `lane == 0`: a lane test makes the region divergent for the compiler, which then wraps every
  // uniform-datapath instruction -- UTMALDG, UTCHMMA, UTCBAR -- in an ELECT / BRA.U.ANY serialisation loop)
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kblocks2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_init(&res_bar, 1);
    fence_mbar_init();
    // The first ring fill needs neither TMEM nor the other warps: issue it now so that the load latency overlaps
    // the TMEM allocation and the CTA-wide barrier below.  (PDL: inputs may only be read after griddepcontrol.wait.
    // Measured: issuing the constant weight tiles before the wait and the activation tiles after it, with or without
    // an early launch_dependents trigger, is slower -- stage 0 completes later and parked successors starve the
    // concurrent streams.)
    pdl_wait();
    for (int kb = kb_begin; kb < kb_prefill; ++kb) produce(kb);
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
  if (threadIdx.x == 0) GEMM_TRACE(1);

  if (warp == 0) {
    // =============================== TMA producer ===========================================
    if (elect_one()) {
      for (int kb = kb_prefill; kb < kb_end; ++kb) produce(kb);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =============================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N, Cvt<T>::kFmt, 0, 0);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int it = kb - kb_begin;
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (it == 0) GEMM_TRACE(2);
        const uint32_t sa = smem_u32(smem + s * kStageBytes);
        const uint32_t sb = sa + kABytes;
        const uint64_t adesc = smem_desc_sw128(sa, 16, 1024);
        const uint64_t bdesc = smem_desc_sw128(sb, 16, 1024);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advancing 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the >>4 address field
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      GEMM_TRACE(3);
      if (has_work) umma_commit(&accum_bar);  // accumulator complete
      else mbar_arrive(&accum_bar);
      pdl_launch_dependents_late();
    }
  } else {
    // =============================== epilogue ===============================================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row of the tile owned by this thread
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
    const bool row_ok = (x < x_end) && (y < p.H) && (img_c < p.NI);
    const long long row = (static_cast<long long>(img_c) * p.H + y) * p.W + x;
    const int img = p.flat ? (p.rows_per_img > 0 ? x / p.rows_per_img : 0) : img_c;

    // A tail tile of a row segment must not store the rows that belong to the next segment: those tiles (and
    // fp32 outputs) take the per-thread store path below; everything else goes through smem + TMA.  The staged
    // per-column vector below needs one image per warp (32 rows).
    const bool rv_uniform = !p.rowvec || (p.flat ? (p.rows_per_img > 0 && p.rows_per_img % 32 == 0)
                                                 : ((p.bw * p.bh) % 32 == 0));
    const bool use_tma_epi = p.tma_epi && !p.coop && rv_uniform && !(p.flat && x0 + kBlockM > x_end && x_end < p.W);
    // ---- while the mainloop runs: fetch this thread's slice of the per-column epilogue vector
    //      vec[warp][col] = bias[col] + rowvec[image of the warp's rows][col]  (registers now, smem after the MMAs)
    constexpr int kVPT = (BLOCK_N + 127) / 128;
    float vpre[4][kVPT];
    const bool ln_in = LN && p.ln_rowstat != nullptr;
    const bool ln_out = LN && p.rowstat_out != nullptr;
    float ln_mean = 0.f, ln_rstd = 1.f;
    if (ln_in && row_ok) {  // LayerNorm statistics of this thread's INPUT row (accumulated by the producing GEMM)
      const float2 sq = *reinterpret_cast<const float2*>(p.ln_rowstat + 2 * row);
      ln_mean = sq.x * p.ln_inv_k;
      ln_rstd = rsqrtf(fmaxf(sq.y * p.ln_inv_k - ln_mean * ln_mean, 0.f) + p.ln_eps);
    }
    if (use_tma_epi) {
      const int et = threadIdx.x - 64;
#pragma unroll
      for (int i = 0; i < kVPT; ++i) {
        const int col = et + 128 * i;
        const bool col_ok = col < BLOCK_N && n0 + col < p.N;
        const float bv = (col_ok && p.bias) ? p.bias[b_noff + n0 + col] : 0.f;
        if (ln_in) {  // slot 0: folded bias, slot 1: column sums of the gamma-scaled weights
          vpre[0][i] = bv;
          vpre[1][i] = col_ok ? p.ln_colsum[b_noff + n0 + col] : 0.f;
          vpre[2][i] = vpre[3][i] = 0.f;
          continue;
        }
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
          float rv = 0.f;
          if (col_ok && p.rowvec) {
            // image of row w4*32 of this tile (the warp's rows all share it)
            const int r4 = w4 * 32;
            int ximg;
            bool ok4;
            if (p.flat) {
              ximg = (x0 + r4) / p.rows_per_img;
              ok4 = x0 + r4 < x_end;
            } else {
              ximg = i0 + r4 / (p.bw * p.bh);
              ok4 = ximg < p.NI;
            }
            if (ok4) rv = p.rowvec[static_cast<long long>(ximg) * p.rowvec_ld + n0 + col];
          }
          vpre[w4][i] = bv + rv;
        }
      }
    }

    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) GEMM_TRACE(4);
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    // ---- split-K: publish this CTA's partial tile; only the last arriver continues to the epilogue ----
    const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
    const long long tiles = static_cast<long long>(gridDim.x) * gridDim.y;
    bool from_ws = false;
    if (p.splits > 1) {
      // layout [split][tile][BLOCK_N/4][128 rows] float4: a warp's 32 rows store 512 contiguous bytes
      float4* wp = reinterpret_cast<float4*>(p.ws_partial) +
                   (static_cast<long long>(blockIdx.z) * tiles + tile_id) * (kBlockM * BLOCK_N / 4) + r;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(t_row + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 f = has_work ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          __stcg(wp + static_cast<long long>(c / 4 + j) * kBlockM, f);
        }
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p.coop) {
        // cooperative reduction: wait until ALL splits of this tile have published (the launcher guarantees that the
        // whole grid is co-resident), then every CTA sums and finishes its own 16-column chunks.  Counter: 0 .. S-1
        // while publishing, S .. 2S-1 while reading, reset by the last reader.
        if (threadIdx.x == 64) {
          atomicAdd(p.ws_counter + tile_id, 1);
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(p.ws_counter + tile_id) < p.splits) {
            __nanosleep(64);
            if (clock64() - t0 > ES_MBAR_TIMEOUT_CYCLES) {
              printf("edgestyle_b200: cooperative split-K timeout tile %d\n", tile_id);
              __trap();
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      } else {
        if (threadIdx.x == 64) {
          const int old = atomicAdd(p.ws_counter + tile_id, 1);
          splitk_last = (old == p.splits - 1) ? 1 : 0;
          if (splitk_last) p.ws_counter[tile_id] = 0;  // self-reset for the next launch
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (!splitk_last) goto epilogue_done;
      }
      __threadfence();
      from_ws = true;
    }
    {
      // accumulator columns [c, c+16) of this thread's row: from TMEM, or the deterministic sum of all partials
      auto load_cols = [&](int c, float (&o)[16]) {
        if (!from_ws) {
          uint32_t v[16];
          tmem_ld_x16(t_row + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = 0.f;
          // three partial tiles in flight per round trip to L2 (the loop was one dependent round trip per split);
          // the summation order stays split 0, 1, 2, ... -> bit-identical, deterministic results
          const long long sstride = tiles * (kBlockM * BLOCK_N / 4);
          const float4* wp0 = reinterpret_cast<const float4*>(p.ws_partial) +
                              static_cast<long long>(tile_id) * (kBlockM * BLOCK_N / 4) +
                              static_cast<long long>(c / 4) * kBlockM + r;
          for (int sp = 0; sp < p.splits; sp += 3) {
            float4 f[3][4];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              const float4* wp = wp0 + static_cast<long long>(sp + u) * sstride;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                f[u][j] = (sp + u < p.splits) ? __ldcg(wp + j * kBlockM) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                o[4 * j] += f[u][j].x; o[4 * j + 1] += f[u][j].y; o[4 * j + 2] += f[u][j].z; o[4 * j + 3] += f[u][j].w;
              }
          }
        }
      };
      auto store16 = [&](T* optr, const float (&o)[16], int valid) {
        if (valid >= 16) {
          uint4 w0, w1;
          w0.x = Cvt<T>::pack2(o[0], o[1]); w0.y = Cvt<T>::pack2(o[2], o[3]);
          w0.z = Cvt<T>::pack2(o[4], o[5]); w0.w = Cvt<T>::pack2(o[6], o[7]);
          w1.x = Cvt<T>::pack2(o[8], o[9]); w1.y = Cvt<T>::pack2(o[10], o[11]);
          w1.z = Cvt<T>::pack2(o[12], o[13]); w1.w = Cvt<T>::pack2(o[14], o[15]);
          reinterpret_cast<uint4*>(optr)[0] = w0;
          reinterpret_cast<uint4*>(optr)[1] = w1;
        } else {
          for (int j = 0; j < 16; ++j)
            if (j < valid) optr[j] = Cvt<T>::from_f(o[j]);
        }
      };

      // GroupNorm statistics of the tile being written (fused producer-side stats: the consumer GroupNorm then
      // only needs its apply pass).  Every lane of a warp holds the same 16 output channels of 32 different pixels
      // of ONE image, so each group segment is warp-reduced and lane 0 issues one atomic per (image, group).
      auto gn_accumulate = [&](const float (&o)[16], int c) {
        int nvalid = p.N - (n0 + c);
        if (nvalid > 16) nvalid = 16;
        if (nvalid <= 0) return;
        const int col0 = p.gn_col0 + n0 + c;  // channel of the normalised tensor (this GEMM may write a column slice)
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
          // rows of a warp never straddle images (checked on the host): use lane 0's image
          const int img0 = __shfl_sync(0xffffffffu, img, 0);
          const int ok0 = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, 0);
          if (lane == 0 && (ok0 || sv != 0.f || qv != 0.f)) {
            float* w = p.gn_ws + (static_cast<long long>(img0) * p.gn_groups + g) * 2;
            atomicAdd(w, sv);
            atomicAdd(w + 1, qv);
          }
        }
      };
      if (use_tma_epi) {
        // ---------------- smem-staged epilogue: 64-column panels [128 rows x 128 B], SWIZZLE_128B ----------------
        // The pipeline's smem is free (every MMA has retired), so the panels alias the stage buffers; the staged
        // epilogue vector sits right behind them.
        constexpr int NOUT = BLOCK_N;                      // accumulator columns per tile
        const bool geglu = p.act == ES_ACT_GEGLU;
        constexpr bool silu_act = SILU;                    // its own instantiation; applied before the residual add
        const int n_tile_out = geglu ? NOUT / 2 : NOUT;    // output columns of this tile
        const int oc0 = geglu ? blockIdx.y * (NOUT / 2) : n0;
        const int full_panels = n_tile_out / 64;
        const int rem = n_tile_out % 64;                   // last, narrower panel (unswizzled, pitch rem*2 B)
        const bool has_res = p.residual != nullptr;
        // (elect.sync, not a thread-index test, around TMA instructions: see the producer warp)
        if (has_res && warp == 2 && elect_one()) {
          mbar_expect_tx(&res_bar, static_cast<uint32_t>(kBlockM * n_tile_out * 2));
          for (int pn = 0; pn < full_panels; ++pn) tma_load_4d(smem + pn * 16384, &tmR, &res_bar, oc0 + pn * 64, x0, y0, i0);
          if (rem) tma_load_4d(smem + full_panels * 16384, &tmRp, &res_bar, oc0 + full_panels * 64, x0, y0, i0);
        }
        float* vec_s = reinterpret_cast<float*>(smem + ((NOUT + 63) / 64) * 16384);  // [4 warps][BLOCK_N]
        float* gstat_s = vec_s + 4 * BLOCK_N;                                        // [4 images][kGnSlots][2]
        // GroupNorm statistics of the finished tile are taken from the smem panels (the rounded values the consumer
        // will read), one 8-column chunk per lane, reduced in smem and flushed with one atomic per (image, group)
        const bool gn_panel = p.gn_ws && !geglu && p.gn_cpg >= 8 && BLOCK_N / p.gn_cpg + 2 <= kGnSlots;
        {
          const int et = threadIdx.x - 64;
          if (gn_panel)
            for (int i = et; i < 4 * kGnSlots * 2; i += 128) gstat_s[i] = 0.f;
#pragma unroll
          for (int i = 0; i < kVPT; ++i) {
            const int col = et + 128 * i;
            if (col < BLOCK_N) {
#pragma unroll
              for (int w4 = 0; w4 < 4; ++w4) vec_s[w4 * BLOCK_N + col] = vpre[w4][i];
            }
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) GEMM_TRACE(9);
        if (has_res) mbar_wait(&res_bar, 0);
        const float* vrow = ln_in ? vec_s : vec_s + q * BLOCK_N;
        const float* srow = vec_s + BLOCK_N;  // LayerNorm-folded GEMM: column sums
        float rs_acc = 0.f, rq_acc = 0.f;     // producer side: (sum, sumsq) of this thread's output row
        constexpr int HALF = NOUT / 2;
        // one 16-column chunk: accumulator (+ gate) -> epilogue math -> this thread's 32 bytes of the smem panel
        auto finish_chunk = [&](int c, float (&o)[16]) {
          const int pn = c >> 6;
          uint8_t* pbase = smem + pn * 16384;
          uint4* d0;
          uint4* d1;
          if (pn < full_panels) {
            const int ch = (c & 63) >> 3;  // 16 B chunk index inside the 128 B row
            d0 = reinterpret_cast<uint4*>(pbase + r * 128 + ((ch ^ (r & 7)) << 4));
            d1 = reinterpret_cast<uint4*>(pbase + r * 128 + (((ch + 1) ^ (r & 7)) << 4));
          } else {
            d0 = reinterpret_cast<uint4*>(pbase + r * (rem * 2) + (c & 63) * 2);
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
          if (p.gn_ws && !geglu && !gn_panel) gn_accumulate(o, c);
          if (ln_out) {
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
        auto vec16 = [&](int col, float (&b)[16]) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 f = *reinterpret_cast<const float4*>(vrow + col + 4 * j);
            b[4 * j] = f.x; b[4 * j + 1] = f.y; b[4 * j + 2] = f.z; b[4 * j + 3] = f.w;
          }
        };
        // accumulator -> pre-activation value: plain bias add, or the folded LayerNorm rstd (acc - mean colsum) + bias'
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
        if (from_ws) {
          // split-K last arriver: the accumulator is the sum of the partial tiles in the workspace
#pragma unroll 1
          for (int c = 0; c < n_tile_out; c += 16) {
            float o[16];
            if (geglu) {
              float a[16], g[16];
              load_cols(c, a);
              load_cols(HALF + c, g);
              pre16(c, a);
              pre16(HALF + c, g);
              epi_geglu16(o, a, g, p.alpha);  // stage by stage across the eight pairs (epilogue.cuh)
            } else {
              load_cols(c, o);
              pre16(c, o);
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = silu_act ? silu_f(o[j] * p.alpha) : o[j] * p.alpha;
            }
            finish_chunk(c, o);
          }
        } else if (geglu) {
          // TMEM loads of chunk c+1 (value and gate columns) are in flight while chunk c is processed
          uint32_t va[2][16], vg[2][16];
          tmem_ld_x16(t_row, va[0]);
          tmem_ld_x16(t_row + HALF, vg[0]);
          auto geglu_chunk = [&](int c, const uint32_t (&a)[16], const uint32_t (&g)[16]) {
            float o[16], av[16], gv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              av[j] = __uint_as_float(a[j]);
              gv[j] = __uint_as_float(g[j]);
            }
            pre16(c, av);
            pre16(HALF + c, gv);
            epi_geglu16(o, av, gv, p.alpha);  // stage by stage across the eight pairs (epilogue.cuh)
            finish_chunk(c, o);
          };
#pragma unroll 1
          for (int c = 0; c < n_tile_out; c += 32) {
            tmem_ld_wait();
            if (c + 16 < n_tile_out) {
              tmem_ld_x16(t_row + c + 16, va[1]);
              tmem_ld_x16(t_row + HALF + c + 16, vg[1]);
            }
            geglu_chunk(c, va[0], vg[0]);
            if (c + 16 < n_tile_out) {
              tmem_ld_wait();
              if (c + 32 < n_tile_out) {
                tmem_ld_x16(t_row + c + 32, va[0]);
                tmem_ld_x16(t_row + HALF + c + 32, vg[0]);
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
            for (int j = 0; j < 16; ++j) o[j] = silu_act ? silu_f(o[j] * p.alpha) : o[j] * p.alpha;
            finish_chunk(c, o);
          };
#pragma unroll 1
          for (int c = 0; c < n_tile_out; c += 32) {
            tmem_ld_wait();
            if (c + 16 < n_tile_out) tmem_ld_x16(t_row + c + 16, v[1]);
            plain_chunk(c, v[0]);
            if (c + 16 < n_tile_out) {
              tmem_ld_wait();
              if (c + 32 < n_tile_out) tmem_ld_x16(t_row + c + 32, v[0]);
              plain_chunk(c + 16, v[1]);
            }
          }
        }
        if (threadIdx.x == 64) GEMM_TRACE(10);
        if (ln_out && row_ok) {
          atomicAdd(p.rowstat_out + 2 * row, rs_acc);
          atomicAdd(p.rowstat_out + 2 * row + 1, rq_acc);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) GEMM_TRACE(11);
        if (warp == 2 && elect_one()) {
          for (int pn = 0; pn < full_panels; ++pn) tma_store_4d(&tmO, smem + pn * 16384, oc0 + pn * 64, x0, y0, i0);
          if (rem) tma_store_4d(&tmOp, smem + full_panels * 16384, oc0 + full_panels * 64, x0, y0, i0);
          tma_store_commit();
        }
        if (gn_panel) {
          const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
          const int img_w = __shfl_sync(0xffffffffu, img, 0);
          const int img_t0 = p.flat ? (p.rows_per_img > 0 ? x0 / p.rows_per_img : 0) : i0;
          const int g_t0 = (p.gn_col0 + n0) / p.gn_cpg;
          const int cc = lane;  // 8-column chunk of the tile row owned by this lane
          const int col0 = n0 + cc * 8;
          if (cc < n_tile_out / 8 && okmask != 0 && col0 < p.N && img_w - img_t0 < 4) {
            float sv[8], qv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j] = qv[j] = 0.f;
            const int pn = cc >> 3;
            const uint8_t* pb = smem + pn * 16384;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              if ((okmask >> rr) & 1u) {
                const int row = q * 32 + rr;
                const uint4 u = (pn < full_panels)
                                    ? *reinterpret_cast<const uint4*>(pb + row * 128 + (((cc & 7) ^ (row & 7)) << 4))
                                    : *reinterpret_cast<const uint4*>(pb + row * (rem * 2) + (cc & 7) * 16);
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
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
            const int bnd = (g0 + 1) * p.gn_cpg - (p.gn_col0 + col0);  // columns [0, bnd) belong to g0, the rest to g0 + 1
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
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int i = threadIdx.x - 64; i < 4 * kGnSlots * 2; i += 128) {
            const float v = gstat_s[i];
            if (v != 0.f) {
              const int il4 = i / (kGnSlots * 2), gl = (i >> 1) % kGnSlots;
              atomicAdd(p.gn_ws + (static_cast<long long>(img_t0 + il4) * p.gn_groups + g_t0 + gl) * 2 + (i & 1), v);
            }
          }
        }
        if (threadIdx.x == 64) GEMM_TRACE(12);
        if (warp == 2 && elect_one()) tma_store_wait_read0();  // (the same thread: elect.sync is deterministic per mask)
      } else if (p.act == ES_ACT_GEGLU) {
        constexpr int HALF = BLOCK_N / 2;
        const int oc0 = blockIdx.y * HALF;  // output column base
        const int n_out = p.N / 2;
#pragma unroll 1
        for (int c = 0; c < HALF; c += 16) {
          float a[16], g[16];
          load_cols(c, a);
          load_cols(HALF + c, g);
          if (row_ok) {
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float av = a[j], gv = g[j];
              if (p.bias) {
                av += p.bias[b_noff + n0 + c + j];
                gv += p.bias[b_noff + n0 + HALF + c + j];
              }
              o[j] = p.alpha * av * gelu_erf_f(gv);
            }
            store16(reinterpret_cast<T*>(p.out) + row * p.ldc + oc0 + c, o, n_out - (oc0 + c));
          }
        }
      } else {
        // software pipeline: the residual (global) and accumulator (TMEM) loads of chunk c+1 are in flight while
        // chunk c is converted and stored
        const bool res_vec = p.residual != nullptr && row_ok;
        const T* rbase = reinterpret_cast<const T*>(p.residual) + (res_vec ? row * p.ldr + n0 : 0);
        uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = make_uint4(0, 0, 0, 0);
        auto fetch_res = [&](int c) {
          if (res_vec && n0 + c + 16 <= p.N) {
            rn0 = reinterpret_cast<const uint4*>(rbase + c)[0];
            rn1 = reinterpret_cast<const uint4*>(rbase + c)[1];
          }
        };
        uint32_t vn[16];
        // cooperative split-K: this CTA finishes chunks [c_lo, c_hi) of the tile, the other splits the rest
        constexpr int kChunks = BLOCK_N / 16;
        const int c_lo = (p.coop && from_ws) ? 16 * ((static_cast<int>(blockIdx.z) * kChunks) / p.splits) : 0;
        const int c_hi = (p.coop && from_ws) ? 16 * (((static_cast<int>(blockIdx.z) + 1) * kChunks) / p.splits) : BLOCK_N;
        fetch_res(c_lo);
        if (!from_ws) {
          tmem_ld_x16(t_row, vn);
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 16) {
          float o[16];
          const uint4 r0 = rn0, r1 = rn1;
          if (!from_ws) {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(vn[j]);
            if (c + 16 < BLOCK_N) tmem_ld_x16(t_row + c + 16, vn);
          } else {
            load_cols(c, o);
          }
          if (c + 16 < c_hi) fetch_res(c + 16);
          if (row_ok && n0 + c < p.N) {
            const int valid = p.N - (n0 + c);
            const bool full = valid >= 16;
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (full || j < valid) o[j] += p.bias[b_noff + n0 + c + j];
            }
            if (p.rowvec) {
              const float* rv = p.rowvec + static_cast<long long>(img) * p.rowvec_ld + n0 + c;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (full || j < valid) o[j] += rv[j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] *= p.alpha;
            if (SILU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = silu_f(o[j]);
            }
            if (p.residual) {
              if (full) {
                const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float2 f = Cvt<T>::unpack2(rr[j]);
                  o[2 * j] += f.x;
                  o[2 * j + 1] += f.y;
                }
              } else {
                const T* rp = reinterpret_cast<const T*>(p.residual) + row * p.ldr + n0 + c;
                for (int j = 0; j < 16; ++j)
                  if (j < valid) o[j] += Cvt<T>::to_f(rp[j]);
              }
            }
            if (p.gn_ws) gn_accumulate(o, c);
            if (p.out_fp32) {
              float* optr = reinterpret_cast<float*>(p.out) + row * p.ldc + n0 + c;
              if (full) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  reinterpret_cast<float4*>(optr)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
              } else {
                for (int j = 0; j < 16; ++j)
                  if (j < valid) optr[j] = o[j];
              }
            } else {
              store16(reinterpret_cast<T*>(p.out) + row * p.ldc + n0 + c, o, valid);
            }
          }
        }
      }
    }
    if (p.coop && p.splits > 1) {  // last reader resets the tile counter for the next launch
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const int old = atomicAdd(p.ws_counter + tile_id, 1);
        if (old == 2 * p.splits - 1) p.ws_counter[tile_id] = 0;
      }
    }
  epilogue_done:;
    if (threadIdx.x == 64) GEMM_TRACE(5);
  }

  // ---- teardown ------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
    if (lane == 0) GEMM_TRACE(6);
  }
}

// --------------------------------------------------------------------------------------------
// host
// --------------------------------------------------------------------------------------------
template <typename T, int BLOCK_N>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2,
                       const CUtensorMap& tmB2, GemmKParams& kp, int m_tiles, int n_tiles, const EsGemm* g,
                       cudaStream_t stream) {
  // ---- epilogue tensor maps: output (and residual) tiles as 64-column panels over the same pixel box as A
  CUtensorMap tmO = tmA, tmOp = tmA, tmR = tmA, tmRp = tmA;
  kp.n_out = g->act == ES_ACT_GEGLU ? g->n / 2 : g->n;
  kp.tma_epi = 0;
  {
    const int n_tile_out = g->act == ES_ACT_GEGLU ? BLOCK_N / 2 : BLOCK_N;
    const int rem = n_tile_out % 64;
    const bool aligned = (reinterpret_cast<uintptr_t>(g->out) & 15) == 0 && g->ldc % 8 == 0 && kp.n_out % 8 == 0 &&
                         (!g->residual || ((reinterpret_cast<uintptr_t>(g->residual) & 15) == 0 && g->ldr % 8 == 0));
    if (!g->out_fp32 && aligned && !getenv("ES_NO_TMA_EPILOGUE")) {
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
      kp.tma_epi = 1;
    }
  }
  if (g->gn_ws && (g->gn_cpg > 0 || g->gn_col0 != 0))
    ES_CHECK(kp.tma_epi, "es_gemm: GroupNorm statistics of a column slice need a 16-bit, 16-byte aligned output");
  if (g->ln_rowstat || g->rowstat_out) {
    // the folded-LayerNorm epilogues exist on the smem/TMA path only
    ES_CHECK(kp.tma_epi, "es_gemm: folded LayerNorm needs a 16-bit, 16-byte aligned output");
    if (kp.flat && kp.nseg > 1)
      for (int sgi = 0; sgi < kp.nseg; ++sgi)
        ES_CHECK((kp.seg_row_start[sgi + 1] - kp.seg_row_start[sgi]) % kBlockM == 0 || sgi == kp.nseg - 1,
                 "es_gemm: folded LayerNorm needs row segments that are multiples of 128 rows");
  }
  constexpr int kStageBytes = kBlockM * kBlockK * 2 + BLOCK_N * kBlockK * 2;
  const int kb_total = kp.taps * kp.kblocks1 + kp.kblocks2;
  const int tiles = m_tiles * n_tiles;
  // ---- split-K: small-M layers have too few output tiles to cover 148 SMs; stream K over several CTAs
  int splits = 1;
  // split_k < 0: COOPERATIVE split-K with -split_k splits -- all split CTAs of a tile wait for each other and each
  // finishes its own column chunks (no serial last-arriver reduction: that one CTA reads splits x 128 x BLOCK_N x 4 B).
  // The CTAs spin on each other, so the whole grid must be co-resident, and so must the grids of two such launches on
  // concurrent streams: at most 148 CTAs of at most half an SM each; only long-K launches qualify (the zero-convs on the
  // third stream never do).
  bool coop = false;
  if (g->workspace && g->split_k < -1) {
    const int want = -g->split_k;
    if (tiles * want <= 148 && kb_total >= 32 && want <= kb_total / 2 && g->act != ES_ACT_GEGLU && !g->ln_rowstat &&
        !g->rowstat_out && BLOCK_N <= 256) {
      coop = true;
      splits = want;
    } else {
      splits = want <= kb_total / 2 ? want : 1;
    }
    const long long need = 65536 + static_cast<long long>(splits) * tiles * kBlockM * BLOCK_N * 4;
    if (splits > 1 && (need > g->workspace_bytes || tiles > 16384)) {
      splits = 1;
      coop = false;
    }
  } else if (g->workspace && g->split_k != 1) {
    if (g->split_k > 1) {
      splits = g->split_k;
    } else if (tiles <= 148 && kb_total >= 16) {
      // fill at most ONE wave of 2 CTAs/SM (a second, nearly empty wave would double the time)
      splits = (2 * 148) / tiles;
      if (splits > kb_total / 6) splits = kb_total / 6;
      if (splits > 32) splits = 32;
      if (splits < 1) splits = 1;
    }
    const long long need = 65536 + static_cast<long long>(splits) * tiles * kBlockM * BLOCK_N * 4;
    if (splits > 1 && (need > g->workspace_bytes || tiles > 16384)) splits = 1;
  }
  kp.splits = splits;
  kp.coop = (coop && splits > 1) ? 1 : 0;
  {
    static int rot_env = -1;
    if (rot_env < 0) {
      const char* e = getenv("ES_K_ROTATE");
      rot_env = e ? atoi(e) : 1;
    }
    kp.k_rotate = (rot_env && m_tiles <= 32 && kb_total >= 16) ? 1 : 0;
  }
  kp.ws_counter = reinterpret_cast<int*>(g->workspace);
  kp.ws_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(g->workspace) + 65536);
  // ---- pipeline depth: default leaves room for two CTAs per SM so one CTA's epilogue overlaps the other's MMAs
  const int kb_cta = (kb_total + splits - 1) / splits;
  // (a deeper ring for single-wave grids was measured: slower in the multi-stream step, because a 200 KB CTA keeps the
  // concurrent streams' CTAs off its SM)
  const int budget = 108 * 1024;
  int stages = g->stages > 0 ? g->stages : budget / kStageBytes;
  if (stages > 12) stages = 12;
  if (stages < 2) stages = 2;
  if (stages > kb_cta) stages = kb_cta < 2 ? 2 : kb_cta;
  kp.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * kStageBytes + 1024;
  dim3 grid(m_tiles, n_tiles, splits);
  // one instantiation per epilogue family, so the common one stays lean (each has its own smem attribute high-water mark)
#define ES_LAUNCH_GEMM(EXT_)                                                                                          \
  do {                                                                                                                \
    auto kern = gemm_kernel<T, BLOCK_N, EXT_>;                                                                        \
    static size_t attr_smem = 0;                                                                                      \
    if (smem > attr_smem) {                                                                                           \
      ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));       \
      attr_smem = smem;                                                                                               \
    }                                                                                                                 \
    ES_CUDA(launch_kernel(kern, dim3(grid), dim3(kGemmThreads), smem, stream, tmA, tmB, tmA2, tmB2, tmO, tmOp, tmR,   \
                          tmRp, kp));                                                                                 \
  } while (0)
  if (g->ln_rowstat || g->rowstat_out) {
    ES_CHECK(g->act != ES_ACT_SILU, "es_gemm: SiLU and folded LayerNorm cannot be combined");
    ES_LAUNCH_GEMM(1);
  } else if (g->act == ES_ACT_SILU) {
    ES_LAUNCH_GEMM(2);
  } else {
    ES_LAUNCH_GEMM(0);
  }
#undef ES_LAUNCH_GEMM
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es
#include "gemm_pair.cuh"
#include "gemm_persist.cuh"
namespace es {

static int pick_block_n(int n, int m_tiles, int act, int kb_total) {
  // Measured on B200 (tools/smallm_sweep.py, tools/gemm_sweep.py):
  //  * small M with a long K (8x8-level convs, M <= 512): 64-wide tiles + split-K stream the weights best;
  //  * mid M (<= 2048 rows) with a long K: the widest tile that divides N, so A is re-read from L2 as little as
  //    possible, again with split-K filling the machine;
  //  * otherwise a wave-quantised cost model over two resident CTAs per SM.
  if (act != ES_ACT_GEGLU && kb_total >= 32) {
    if (m_tiles <= 4 && n >= 64) return 64;
    if (m_tiles <= 16) {
      if (n % 160 == 0) return 160;
      if (n % 128 == 0) return 128;
    }
  }
  const int cands[5] = {256, 160, 128, 64, 32};
  int best = 32;
  double best_cost = 1e30;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (act == ES_ACT_GEGLU && (bn % 32 != 0)) continue;
    const int nt = ceil_div(n, bn);
    const double waste = static_cast<double>(nt) * bn / n;  // >= 1
    const int ctas = nt * m_tiles;
    const int waves = ceil_div(ctas, 296);  // two CTAs per SM
    // time ~ waves * (per-tile time ~ bn + fixed overhead)
    const double cost = waves * (bn + 48.0) * (0.9 + 0.1 * waste);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

template <typename T>
static int gemm_dispatch(const EsGemm* g, cudaStream_t stream) {
  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const bool flat = (g->taps == 1);
  ES_CHECK(g->taps == 1 || g->taps == 9, "es_gemm: taps must be 1 or 9 (got %d)", g->taps);
  ES_CHECK(g->c1 > 0 && g->c1 % 8 == 0 && g->lda % 8 == 0, "es_gemm: c1/lda must be multiples of 8 (c1=%d lda=%lld)",
           g->c1, g->lda);
  ES_CHECK((reinterpret_cast<uintptr_t>(g->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->b) & 15) == 0,
           "es_gemm: a/b must be 16-byte aligned");
  ES_CHECK(g->n > 0 && g->w > 0 && g->h > 0 && g->n_img > 0, "es_gemm: bad shape");
  ES_CHECK(g->ldc % 8 == 0 || g->out_fp32, "es_gemm: ldc must be a multiple of 8");
  kp.flat = flat ? 1 : 0;
  kp.N = g->n;
  kp.taps = g->taps;
  kp.kblocks1 = ceil_div(g->c1, kBlockK);
  kp.kblocks2 = 0;
  int m_tiles;
  int W, H, NI;
  if (flat) {
    // treat [n_img, h, w] as one flat row dimension
    const long long M = static_cast<long long>(g->w) * g->h * g->n_img;
    ES_CHECK(M < (1ll << 31), "es_gemm: M too large");
    W = static_cast<int>(M);
    H = 1;
    NI = 1;
    kp.bw = kBlockM;
    kp.bh = 1;
    kp.bn = 1;
    kp.nseg = g->nseg > 0 ? g->nseg : 1;
    ES_CHECK(kp.nseg <= ES_MAX_SEG, "es_gemm: too many segments");
    int tiles = 0;
    for (int s = 0; s < kp.nseg; ++s) {
      const int r0 = g->nseg > 0 ? g->seg_row_start[s] : 0;
      const int r1 = g->nseg > 0 ? g->seg_row_start[s + 1] : W;
      ES_CHECK(r1 >= r0 && r1 <= W, "es_gemm: bad segment bounds");
      kp.seg_row_start[s] = r0;
      kp.seg_row_start[s + 1] = r1;
      kp.seg_tile_start[s] = tiles;
      tiles += ceil_div(r1 - r0, kBlockM);
      kp.seg_b_noff[s] = g->nseg > 0 ? g->seg_b_noff[s] : 0;
      kp.seg_b2_noff[s] = g->a2 ? (g->nseg > 0 ? g->seg_b2_noff[s] : 0) : -1;
    }
    kp.seg_tile_start[kp.nseg] = tiles;
    m_tiles = tiles;
    kp.tiles_x = tiles;
    kp.tiles_y = 1;
  } else {
    W = g->w;
    H = g->h;
    NI = g->n_img;
    int bw = W >= kBlockM ? kBlockM : W;
    ES_CHECK(kBlockM % bw == 0, "es_gemm: conv width %d must divide 128 or be >= 128", W);
    int bh = kBlockM / bw;
    int bn = 1;
    if (bh > H) {
      // several whole images per tile (8x8 level): H must divide bh
      int hh = 1;
      while (hh < H) hh <<= 1;  // next pow2 >= H
      bn = bh / hh;
      bh = hh;
    }
    kp.bw = bw;
    kp.bh = bh;
    kp.bn = bn;
    kp.tiles_x = ceil_div(W, bw);
    kp.tiles_y = ceil_div(H, bh);
    m_tiles = kp.tiles_x * kp.tiles_y * ceil_div(NI, bn);
    kp.nseg = g->nseg > 0 ? g->nseg : 1;
    ES_CHECK(kp.nseg <= ES_MAX_SEG, "es_gemm: too many segments");
    for (int s = 0; s < kp.nseg; ++s) {  // image segments: every boundary must fall on a tile boundary
      const int i0s = g->nseg > 0 ? g->seg_row_start[s] : 0;
      const int i1s = g->nseg > 0 ? g->seg_row_start[s + 1] : NI;
      ES_CHECK(i1s >= i0s && i1s <= NI && (i0s % bn == 0 || i0s == i1s),
               "es_gemm: image segment [%d, %d) does not fall on tiles of %d image(s)", i0s, i1s, bn);
      kp.seg_row_start[s] = i0s;
      kp.seg_row_start[s + 1] = i1s;
      kp.seg_b_noff[s] = g->nseg > 0 ? g->seg_b_noff[s] : 0;
      kp.seg_b2_noff[s] = g->a2 ? (g->nseg > 0 ? g->seg_b2_noff[s] : 0) : -1;
    }
  }
  kp.W = W;
  kp.H = H;
  kp.NI = NI;
  kp.bias = g->bias;
  kp.rowvec = g->rowvec;
  kp.rows_per_img = g->rows_per_img;
  kp.rowvec_ld = g->rowvec_ld;
  kp.residual = g->residual;
  kp.ldr = g->ldr;
  kp.act = g->act;
  kp.alpha = g->alpha;
  kp.out = g->out;
  kp.ldc = g->ldc;
  kp.out_fp32 = g->out_fp32;
  if (g->residual) ES_CHECK(g->ldr % 8 == 0, "es_gemm: ldr must be a multiple of 8");
  kp.prefetch = g->prefetch;
  kp.prefetch_bytes = g->prefetch_bytes;
  kp.rowstat_out = g->rowstat_out;
  kp.ln_rowstat = g->ln_rowstat;
  kp.ln_colsum = g->ln_colsum;
  kp.ln_inv_k = g->ln_features > 0 ? 1.0f / static_cast<float>(g->ln_features) : 0.f;
  kp.ln_eps = g->ln_eps;
  if (g->ln_rowstat) ES_CHECK(g->ln_colsum && g->ln_features > 0 && !g->rowvec, "es_gemm: bad folded-LayerNorm config");
  kp.gn_ws = g->gn_ws;
  kp.gn_groups = g->gn_groups;
  kp.gn_cpg = g->gn_cpg > 0 ? g->gn_cpg : (g->gn_groups > 0 ? g->n / g->gn_groups : 0);
  kp.gn_col0 = g->gn_col0;
  if (g->gn_ws) {
    ES_CHECK(g->gn_groups > 0 && g->act == ES_ACT_NONE, "es_gemm: bad fused-GroupNorm config");
    if (g->gn_cpg > 0 || g->gn_col0 != 0)  // this GEMM writes a column slice of the normalised tensor (concat half)
      ES_CHECK(g->gn_cpg >= 8 && g->gn_col0 >= 0 && (g->gn_col0 + g->n + g->gn_cpg - 1) / g->gn_cpg <= g->gn_groups,
               "es_gemm: bad GroupNorm slice (cpg %d, first channel %d, n %d, groups %d)", g->gn_cpg, g->gn_col0, g->n,
               g->gn_groups);
    else
      ES_CHECK(g->n % g->gn_groups == 0, "es_gemm: n %% gn_groups != 0");
    // a warp's 32 tile rows must belong to one image
    const long long rpi = flat ? g->rows_per_img : static_cast<long long>(g->w) * g->h;
    ES_CHECK(rpi > 0 && rpi % 32 == 0, "es_gemm: fused GroupNorm statistics need rows-per-image %% 32 == 0 (got %lld)", rpi);
    if (flat) ES_CHECK(g->nseg <= 1 || true, "es_gemm: ok");
  }
  if (g->act == ES_ACT_GEGLU) ES_CHECK(g->n % 2 == 0 && !g->out_fp32 && !g->residual && !g->rowvec, "es_gemm: bad GEGLU config");
  ES_CHECK(g->act == ES_ACT_NONE || g->act == ES_ACT_GEGLU || g->act == ES_ACT_SILU, "es_gemm: unknown activation %d", g->act);

  int bn_tile = g->block_n > 0 ? g->block_n : pick_block_n(g->n, m_tiles, g->act, g->taps * ceil_div(g->c1, kBlockK));
  // block_n == 1000 + BN selects the persistent kernel (one CTA per SM, two accumulators in TMEM) with BN-wide tiles;
  // problems it cannot run fall back to the one-tile kernel of the same width
  int persist = 0;
  if (bn_tile >= 1000) {
    const int bn = bn_tile - 1000;
    ES_CHECK(bn == 128 || bn == 160 || bn == 256, "es_gemm: unsupported persistent tile width %d", bn);
    const bool ok = bn == 128 ? gemm_persist_eligible<128>(g, kp) : bn == 160 ? gemm_persist_eligible<160>(g, kp)
                                                                             : gemm_persist_eligible<256>(g, kp);
    if (ok) persist = bn;
    bn_tile = bn;
  }
  // block_n == 320 selects the CTA-pair kernel (256 x 320 tiles, cta_group::2); problems it cannot run fall back
  bool pair = false;
  if (bn_tile == kPairN) {
    pair = gemm_pair_eligible(g, kp, m_tiles);
    if (!pair) bn_tile = g->act == ES_ACT_GEGLU ? kPairNH : pick_block_n(g->n, m_tiles, g->act, g->taps * ceil_div(g->c1, kBlockK));
  }
  const int n_tiles = ceil_div(g->n, bn_tile);
  if (g->act == ES_ACT_GEGLU) ES_CHECK(g->n % bn_tile == 0, "es_gemm: GEGLU needs n %% block_n == 0 (n=%d bn=%d)", g->n, bn_tile);

  // ---- tensor maps ----------------------------------------------------------------------
  CUtensorMap tmA, tmB, tmA2, tmB2;
  {
    const uint64_t pitch = static_cast<uint64_t>(g->lda) * 2;
    uint64_t dims[4], strides[4];
    uint32_t box[4] = {static_cast<uint32_t>(kBlockK), static_cast<uint32_t>(kp.bw), static_cast<uint32_t>(kp.bh),
                       static_cast<uint32_t>(kp.bn)};
    dims[0] = g->c1;
    dims[1] = W;
    dims[2] = H;
    dims[3] = NI;
    strides[1] = pitch;
    strides[2] = pitch * W;
    strides[3] = pitch * W * H;
    if (encode_tmap_16b(&tmA, g->a, 4, dims, strides, box)) return -3;
  }
  kp.b_blocked = g->b_blocked;
  if (pair || persist) {
    tmB = tmA;  // the pair / persistent launchers build their own weight maps
  } else if (g->b_blocked) {
    // K-block-major weights: [taps * kblocks1][n_total_b][64] -- every B tile is one contiguous BLOCK_N x 128 B chunk
    uint64_t dims[3] = {64, static_cast<uint64_t>(g->n_total_b), static_cast<uint64_t>(g->taps) * kp.kblocks1};
    uint64_t strides[3] = {0, 128, static_cast<uint64_t>(g->n_total_b) * 128};
    uint32_t box[3] = {static_cast<uint32_t>(kBlockK), static_cast<uint32_t>(bn_tile), 1u};
    ES_CHECK(g->n_total_b >= g->n, "es_gemm: n_total_b < n");
    if (encode_tmap_16b(&tmB, g->b, 3, dims, strides, box)) return -3;
  } else {
    uint64_t dims[3] = {static_cast<uint64_t>(g->c1), static_cast<uint64_t>(g->taps),
                        static_cast<uint64_t>(g->n_total_b)};
    uint64_t strides[3] = {0, static_cast<uint64_t>(g->c1) * 2, static_cast<uint64_t>(g->c1) * 2 * g->taps};
    uint32_t box[3] = {static_cast<uint32_t>(kBlockK), 1u, static_cast<uint32_t>(bn_tile)};
    ES_CHECK(g->n_total_b >= g->n, "es_gemm: n_total_b < n");
    if (encode_tmap_16b(&tmB, g->b, 3, dims, strides, box)) return -3;
  }
  tmA2 = tmA;
  tmB2 = tmB;
  if (g->a2) {
    ES_CHECK(g->c2 > 0 && g->c2 % 8 == 0 && g->lda2 % 8 == 0 && g->b2, "es_gemm: bad source 2");
    kp.kblocks2 = ceil_div(g->c2, kBlockK);
    const uint64_t pitch = static_cast<uint64_t>(g->lda2) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(g->c2), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(NI)};
    uint64_t strides[4] = {0, pitch, pitch * W, pitch * W * H};
    uint32_t box[4] = {static_cast<uint32_t>(kBlockK), static_cast<uint32_t>(kp.bw), static_cast<uint32_t>(kp.bh),
                       static_cast<uint32_t>(kp.bn)};
    if (encode_tmap_16b(&tmA2, g->a2, 4, dims, strides, box)) return -3;
    uint64_t dimsb[3] = {static_cast<uint64_t>(g->c2), 1, static_cast<uint64_t>(g->n_total_b2)};
    uint64_t stridesb[3] = {0, static_cast<uint64_t>(g->c2) * 2, static_cast<uint64_t>(g->c2) * 2};
    uint32_t boxb[3] = {static_cast<uint32_t>(kBlockK), 1u, static_cast<uint32_t>(bn_tile)};
    if (!pair && !persist && encode_tmap_16b(&tmB2, g->b2, 3, dimsb, stridesb, boxb)) return -3;
  }
  if (persist) {
    ES_CHECK(g->n_total_b >= g->n, "es_gemm: n_total_b < n");
    if (persist == 128) return launch_gemm_persist<T, 128>(tmA, tmA2, kp, m_tiles, g, stream);
    if (persist == 160) return launch_gemm_persist<T, 160>(tmA, tmA2, kp, m_tiles, g, stream);
    return launch_gemm_persist<T, 256>(tmA, tmA2, kp, m_tiles, g, stream);
  }
  if (pair) {
    ES_CHECK(g->n_total_b >= g->n, "es_gemm: n_total_b < n");
    return launch_gemm_pair<T>(tmA, tmA2, kp, m_tiles, g, stream);
  }

  switch (bn_tile) {
    case 32: return launch_gemm<T, 32>(tmA, tmB, tmA2, tmB2, kp, m_tiles, n_tiles, g, stream);
    case 64: return launch_gemm<T, 64>(tmA, tmB, tmA2, tmB2, kp, m_tiles, n_tiles, g, stream);
    case 128: return launch_gemm<T, 128>(tmA, tmB, tmA2, tmB2, kp, m_tiles, n_tiles, g, stream);
    case 160: return launch_gemm<T, 160>(tmA, tmB, tmA2, tmB2, kp, m_tiles, n_tiles, g, stream);
    case 256: return launch_gemm<T, 256>(tmA, tmB, tmA2, tmB2, kp, m_tiles, n_tiles, g, stream);
    default: ES_CHECK(false, "es_gemm: unsupported block_n %d", bn_tile);
  }
  return 0;
}

}  // namespace es

#ifdef ES_GEMM_TRACE
extern "C" int es_gemm_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, es::g_gemm_trace, sizeof(es::g_gemm_trace)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int es_gemm(const EsGemm* g, void* stream) {
  if (!g) {
    es::set_error("es_gemm: null descriptor");
    return -1;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (g->dtype == ES_DTYPE_F16) return es::gemm_dispatch<__half>(g, s);
  if (g->dtype == ES_DTYPE_BF16) return es::gemm_dispatch<__nv_bfloat16>(g, s);
  es::set_error("es_gemm: unknown dtype %d", g->dtype);
  return -1;
}
