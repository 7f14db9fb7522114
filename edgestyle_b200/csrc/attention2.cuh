// es_attention, second generation: QT query tiles (128 rows each) per CTA, one softmax warpgroup per tile, one TMA
// warp and one MMA warp per CTA.  What changed against attention_kernel (attention.cu) and why:
//
//   * the MUFU pipe was 50 % busy with 2 softmax warps per SM sub-partition (one warpgroup per CTA, two CTAs per SM):
//     every warp sat half of its time in the tcgen05.ld -> ex2 -> pack -> st.shared -> fence -> mbarrier chain.  Two
//     query tiles per CTA (K/V tiles shared, loaded once) and still two CTAs per SM give FOUR softmax warps per
//     sub-partition to hide that chain.
//   * S is single-buffered per tile (64 fp32 columns) and released EARLY: as soon as a thread has pulled its row into
//     registers (and checked the row max against the current shift) the warpgroup arrives on s_free and the MMA warp
//     issues Q K^T of the next key tile into the same columns, while the exponentials of the current one run.
//     (Consuming the row in four 16-column chunks with the release after the last load measured slower: 491 vs 440 us.)
//   * one exponential per score on the MUFU pipe in fp32 (ex2.approx.ftz.f32).  Measured on B200
//     (tools/ubench_softmax.cu, profiles/r2r_ubench_softmax.txt): ex2.approx.f16x2 is NOT faster -- 16 packed
//     instructions take the 256 cycles of 32 scalar ones (it is two MUFU operations plus pack/unpack moves).  Pulling
//     the fp32 S tile out of TMEM is cheap (tcgen05.ld 32x32b.x32: 20-38 cycles per warp, overlapping the exponentials);
//     the floor is the MUFU pipe, 512 cycles per 128 x 64 tile and SM sub-partition, and what a warp loses on top is its
//     own dependency chain (load -> max -> exp -> pack -> st.shared -> fence -> mbarrier) -- hence four warps per
//     sub-partition.
//   * P is single-buffered per tile: the write of P(j+1) waits for the commit of P(j) V (pv_done), which also is the
//     "O is stable" condition the rare rescale path needs.
//
//   TMEM columns: [t * 64, +64) = S of tile t, [QT * 64 + t * o_stride, +dN) = O of tile t, o_stride = dN rounded up to
//   32 columns (d = 40: 2 x 64 + 2 x 64 = 256 columns, two CTAs per SM).
#pragma once

namespace es {

constexpr int kAtt2KV = 64;  // keys per K/V tile
#ifndef ES_ATT2_POLY_MASK
#define ES_ATT2_POLY_MASK 0x1084
#endif
// which of the 16 score pairs of a 32-score half row take the FMA-pipe exponential (bit p = pair p); 0 = none
constexpr uint32_t kAtt2PolyMask = ES_ATT2_POLY_MASK;
#ifndef ES_ATT2_STAGGER
#define ES_ATT2_STAGGER 1  // 1: the two query tiles of a CTA run half a key tile apart (see the MMA issuer)
#endif

template <typename T, int NA, int QT>
__global__ void __launch_bounds__(64 + 128 * QT, (NA == 1) ? 2 : 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  constexpr int kKV = kAtt2KV;
  constexpr int kKVAtom = kKV * 128;  // bytes of one 64-column atom of a K / V tile
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                              // QT tiles x NA atoms x [128 rows x 128 B]
  uint8_t* sK = sQ + QT * NA * kAtomBytes;         // 2 stages x NA atoms x [64 rows x 128 B]
  uint8_t* sV = sK + 2 * NA * kKVAtom;             // 2 stages x NA atoms x [64 rows x 128 B]
  uint8_t* sP = sV + 2 * NA * kKVAtom;             // QT x [128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + QT * kAtomBytes);
  uint64_t& q_full = bars[0];
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [QT] MMA -> softmax: S(j) complete
  uint64_t* s_free = bars + 11;   // [QT] softmax -> MMA: S(j) is in registers (128 arrivals)
  uint64_t* p_full = bars + 13;   // [QT] softmax -> MMA: P(j) in smem (128 arrivals)
  uint64_t* pv_done = bars + 15;  // [QT] MMA -> softmax: P(j) V retired (P buffer free, O stable)
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 17);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (128 * QT);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.nkv + kKV - 1) / kKV;
  const uint32_t tmem_cols = p.tmem_cols;
  const int o_col0 = QT * kKV;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < QT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // PDL: everything above overlaps the previous kernel's tail

  if (warp == 0) {
    // ================================ TMA producer ==============================================================
    if (elect_one()) {
      mbar_expect_tx(&q_full, QT * NA * kAtomBytes);
#pragma unroll
      for (int t = 0; t < QT; ++t)
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_4d(sQ + (t * NA + a) * kAtomBytes, &tmQ, &q_full, a * 64, head, q0 + t * 128, b);
      // K runs one key tile ahead of V: K(j+1) is wanted right after S(j) has been pulled, V(j) only once P(j) is in smem,
      // and a V stage is held until the LATER of the two query tiles has issued its P V
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], NA * kKVAtom);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_4d(sK + (s * NA + a) * kKVAtom, &tmK, &k_full[s], a * 64, head, j * kKV, b);
      };
      load_k(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        if (j + 1 < n_tiles) load_k(j + 1);
        mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], NA * kKVAtom);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_4d(sV + (s * NA + a) * kKVAtom, &tmV, &v_full[s], a * 64, head, j * kKV, b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================================================
    // The issuing thread's own instruction stream is the critical resource (measured, tools/att_trace.py: with shared-
    // memory descriptors rebuilt per MMA the thread needed ~3000 cycles per key tile for its 14 MMAs and was what every
    // softmax warpgroup waited for).  Everything loop-invariant is therefore hoisted: a descriptor is its constant high
    // word plus a low word that is one add away from a precomputed base.  The thread is chosen with elect.sync, not
    // `lane == 0`: under a lane test the compiler must assume a divergent region and wraps EVERY uniform-datapath
    // instruction (UTCHMMA, UTCBAR, R2UR) in an ELECT / BRA.U.ANY serialisation loop (~95 cycles per tcgen05.mma).
    if (elect_one()) {
      const uint32_t idesc_qk = make_idesc_f16(128, kKV, Cvt<T>::kFmt, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, p.dN, Cvt<T>::kFmt, 0, 1);
      const int kq = (p.d + 15) / 16;  // MMAs along the head dim
      const uint32_t hi_kmaj = static_cast<uint32_t>(smem_desc_sw128(0, 16, 1024) >> 32);
      const uint32_t hi_v = static_cast<uint32_t>(smem_desc_sw128(0, kKVAtom, 1024) >> 32);
      auto lo_of = [](uint32_t saddr, uint32_t lbo) {
        return static_cast<uint32_t>(smem_desc_sw128(saddr, lbo, 1024) & 0xffffffffull);
      };
      auto mk = [](uint32_t lo, uint32_t hi) { return (static_cast<uint64_t>(hi) << 32) | lo; };
      uint32_t q_lo[QT][NA], k_lo0[NA], p_lo[QT];
#pragma unroll
      for (int t = 0; t < QT; ++t) {
        p_lo[t] = lo_of(smem_u32(sP) + t * kAtomBytes, 16);
#pragma unroll
        for (int a = 0; a < NA; ++a) q_lo[t][a] = lo_of(smem_u32(sQ) + (t * NA + a) * kAtomBytes, 16);
      }
#pragma unroll
      for (int a = 0; a < NA; ++a) k_lo0[a] = lo_of(smem_u32(sK) + a * kKVAtom, 16);
      const uint32_t v_lo0 = lo_of(smem_u32(sV), kKVAtom);
      constexpr uint32_t kStageLo = (NA * kKVAtom) >> 4;  // low-word distance between the two K (or V) stages
      // a second tile that lies entirely beyond nq still runs (its rows are never stored): uniform control flow
      auto issue_qk = [&](int t, int st) {
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            if (a * 4 + k4 < kq)
              umma_f16(tmem_base + t * kKV, mk(q_lo[t][a] + 2 * k4, hi_kmaj), mk(k_lo0[a] + st * kStageLo + 2 * k4, hi_kmaj),
                       idesc_qk, (a | k4) != 0);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {  // O(t) (+)= P(j) V(j); V: MN-major, 16 keys = 2048 B along K; N atoms kKVAtom apart
        const uint32_t v_lo = v_lo0 + (j & 1) * kStageLo;
#pragma unroll
        for (int k = 0; k < kKV / 16; ++k)
          umma_f16(tmem_base + o_col0 + t * p.o_stride, mk(p_lo[t] + 2 * k, hi_kmaj), mk(v_lo + k * (2048 >> 4), hi_v),
                   idesc_pv, (j | k) != 0);
        umma_commit(&pv_done[t]);
      };
      mbar_wait(&q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if constexpr (QT == 2 && ES_ATT2_STAGGER) {
        // The two query tiles run HALF A KEY TILE APART.  In phase, the four softmax warps of a sub-partition (two tiles
        // x two CTAs) all pull S, then all want the MUFU pipe, then all store P: a period of ~500 + 4 x 416 + 450
        // cycles with the pipe idle through the first and last part (tools/att_trace.py).  Tile 1 therefore starts
        // after tile 0 has released S(0), and the issue order below -- QK0(j+1), PV1(j-1), QK1(j+1), PV0(j), each
        // behind its own barrier -- keeps it there: one tile's exponentials run under the other's TMEM / smem phases.
        // (Also starting the SM's second CTA a quarter period late -- four phases per sub-partition -- changes nothing:
        // 352 vs 354 us.)
        issue_qk(0, 0);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j & 1;
          if (j + 1 < n_tiles) {  // [A] S0(j+1) as soon as tile 0 has pulled S0(j)
            mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1);
            mbar_wait(&s_free[0], j & 1);
            tc_fence_after();
            issue_qk(0, s ^ 1);
          }
          ATT_TRACE(0, 4 * j + 0);
          if (j == 0) {           // [B] tile 1 starts here ...
            issue_qk(1, 0);
            umma_commit(&k_empty[0]);
          } else {                //     ... and stays one P V behind
            mbar_wait(&p_full[1], (j - 1) & 1);
            tc_fence_after();
            issue_pv(1, j - 1);
            umma_commit(&v_empty[s ^ 1]);
          }
          ATT_TRACE(0, 4 * j + 1);
          if (j + 1 < n_tiles) {  // [C] S1(j+1)
            mbar_wait(&s_free[1], j & 1);
            tc_fence_after();
            issue_qk(1, s ^ 1);
            umma_commit(&k_empty[s ^ 1]);
          }
          ATT_TRACE(0, 4 * j + 2);
          mbar_wait(&v_full[s], (j >> 1) & 1);  // [D] O0 += P0(j) V(j)
          mbar_wait(&p_full[0], j & 1);
          tc_fence_after();
          issue_pv(0, j);
          ATT_TRACE(0, 4 * j + 3);
        }
        mbar_wait(&p_full[1], (n_tiles - 1) & 1);
        tc_fence_after();
        issue_pv(1, n_tiles - 1);
        umma_commit(&v_empty[(n_tiles - 1) & 1]);
      } else {
#pragma unroll
        for (int t = 0; t < QT; ++t) issue_qk(t, 0);
        umma_commit(&k_empty[0]);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j & 1;
          const uint32_t ph = (j >> 1) & 1;
          if (j + 1 < n_tiles) {  // S(j+1) = Q K(j+1)^T as soon as S(j) has been pulled into registers
            mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1);
#pragma unroll
            for (int t = 0; t < QT; ++t) {
              mbar_wait(&s_free[t], j & 1);
              tc_fence_after();
              issue_qk(t, s ^ 1);
            }
            umma_commit(&k_empty[s ^ 1]);
          }
          ATT_TRACE(0, 4 * j + 0);  // S(j) of both tiles consumed, Q K(j+1)^T issued
          mbar_wait(&v_full[s], ph);
          ATT_TRACE(0, 4 * j + 1);
#pragma unroll
          for (int t = 0; t < QT; ++t) {
            mbar_wait(&p_full[t], j & 1);  // P(j) of tile t in smem
            tc_fence_after();
            issue_pv(t, j);
            ATT_TRACE(0, 4 * j + 2 + (t ? 1 : 0));  // P(j) V of tile t issued
          }
          umma_commit(&v_empty[s]);
        }
      }
    }
  } else {
    // ================================ softmax: warpgroup t owns query tile t, thread r its row r ===============
    const int t = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const uint32_t t_s = t_row + t * kKV;
    const uint32_t t_o = t_row + o_col0 + t * p.o_stride;
    float m_run = -INFINITY;  // shift of the exponent (scaled, log2 units): within 2^kSlack of the running row max
    float l_run = 0.f;        // running row sum of P (fp32)
    const float sl2 = p.scale_log2;
    constexpr float kSlack = 8.0f;
    uint8_t* prow = sP + t * kAtomBytes + r * 128;

    for (int j = 0; j < n_tiles; ++j) {
      const int kv_valid = min(kKV, p.nkv - j * kKV);
      const bool full_tile = kv_valid == kKV;
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      if (lane == 0 && (j == 8 || j == 9)) ATT_TRACE(j - 7, (warp - 2) * 4 + 0);
      uint32_t va[32], vb[32];
      tmem_ld_x32(t_s, va);
      tmem_ld_x32(t_s + 32, vb);
      tmem_ld_wait();
      // row max of the tile (raw scores); a partial last tile masks its tail
      float mx = -INFINITY;
      if (full_tile) {
        // four independent chains: one chain of 32 dependent FMNMX3 is ~190 cycles in which the warp issues nothing else
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            m4[c] = fmax3(m4[c], __uint_as_float(va[i + 2 * c]), __uint_as_float(va[i + 2 * c + 1]));
            m4[c] = fmax3(m4[c], __uint_as_float(vb[i + 2 * c]), __uint_as_float(vb[i + 2 * c + 1]));
          }
        }
        mx = fmaxf(fmax3(m4[0], m4[1], m4[2]), m4[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) va[i] = 0xff800000u;       // -inf: exp -> 0, max ignores it
          if (32 + i >= kv_valid) vb[i] = 0xff800000u;
          mx = fmaxf(mx, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
        }
      }
      bool pv_waited = false;
      const float m_tile = mx * sl2;
      if (j == 0) {
        m_run = m_tile;  // first tile: its own max is the shift
      } else if (__any_sync(0xffffffffu, m_tile > m_run + kSlack)) {
        // rare path: raise the shift of the rows that need it and rescale their O (softmax is shift-invariant: the
        // shift only has to stay within 2^kSlack of the running max, so that P <= 2^kSlack fits 16-bit floats)
        mbar_wait(&pv_done[t], (j - 1) & 1);  // P(j-1) V retired: O is stable (and P is free)
        pv_waited = true;
        tc_fence_after();
        const float m_new = fmaxf(m_run, m_tile);
        const float alpha = ex2_approx(m_run - m_new);
#pragma unroll 1
        for (int c = 0; c < p.dN; c += 16) {
          uint32_t o[16];
          tmem_ld_x16(t_o + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(t_o + c, o);
        }
        tmem_st_wait();
        l_run *= alpha;
        m_run = m_new;
      }
      // S(j) is in registers: the MMA warp may overwrite it with S(j+1) while the exponentials run
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      if (lane == 0 && (j == 8 || j == 9)) ATT_TRACE(j - 7, (warp - 2) * 4 + 1);
      const float neg_m = -m_run;
      uint64_t sum2 = pk2(0.f, 0.f);
      // exponentials: the score pairs selected by kAtt2PolyMask go through a polynomial on the FMA pipe (packed fp32
      // pairs: Cody-Waite split with the 1.5 * 2^23 rounding constant, degree-3 minimax of 2^f on [-0.5, 0.5], relative
      // error 7.5e-5 < half an ulp of the 16-bit P it feeds), the rest through MUFU ex2 -- the MUFU pipe (16 ex2 / clk / SM)
      // is this kernel's floor, the FMA pipe has 8 x its rate
      const uint64_t sl2_2 = pk2(sl2, sl2), negm_2 = pk2(neg_m, neg_m);
      auto exp_pair = [&](uint32_t a, uint32_t b, bool poly, float& p0, float& p1) {
        float x0, x1;
        upk2(fma2(pk2(__uint_as_float(a), __uint_as_float(b)), sl2_2, negm_2), x0, x1);
        if (!poly) {
          p0 = ex2_approx(x0);
          p1 = ex2_approx(x1);
          return;
        }
        const uint64_t x2 = pk2(fmaxf(x0, -120.f), fmaxf(x1, -120.f));  // (-inf of a masked tail included)
        const uint64_t t2 = add2(x2, pk2(12582912.f, 12582912.f));      // n = round(x) in the low mantissa bits
        const uint64_t n2 = add2(t2, pk2(-12582912.f, -12582912.f));
        const uint64_t f2 = fma2(n2, pk2(-1.f, -1.f), x2);              // f = x - n in [-0.5, 0.5]
        uint64_t q2 = fma2(f2, pk2(0.0551716685f, 0.0551716685f), pk2(0.2426111251f, 0.2426111251f));
        q2 = fma2(q2, f2, pk2(0.6932609677f, 0.6932609677f));
        q2 = fma2(q2, f2, pk2(0.9999280572f, 0.9999280572f));
        float q0, q1, t0, t1;
        upk2(q2, q0, q1);
        upk2(t2, t0, t1);
        p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));  // q * 2^n
        p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
      };
      // P(j) goes to smem chunk by chunk as it is produced (canonical K-major SWIZZLE_128B: row r at r * 128 B, 16 B
      // chunk index XOR (r & 7)): the eight 16-byte stores run under the exponentials instead of after them (the
      // separate store phase was ~300 of the ~2400 cycles of a key tile).  The MMAs reading P(j-1) must have retired.
      if (j > 0 && !pv_waited) mbar_wait(&pv_done[t], (j - 1) & 1);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (q4 & 3) * 8 + 2 * e;  // element pair inside the 32-score half row
          float p0, p1;
          if (q4 < 4) exp_pair(va[i], va[i + 1], (kAtt2PolyMask >> (i >> 1)) & 1u, p0, p1);
          else exp_pair(vb[i], vb[i + 1], (kAtt2PolyMask >> (i >> 1)) & 1u, p0, p1);
          sum2 = add2(sum2, pk2(p0, p1));
          w[e] = Cvt<T>::pack2(p0, p1);
        }
        const int chunk = q4 ^ (r & 7);
        *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      float s0, s1;
      upk2(sum2, s0, s1);
      l_run += s0 + s1;
      if (lane == 0 && (j == 8 || j == 9)) ATT_TRACE(j - 7, (warp - 2) * 4 + 2);
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(&p_full[t]);
      if (lane == 0 && (j == 8 || j == 9)) ATT_TRACE(j - 7, (warp - 2) * 4 + 3);
    }
    // ---- epilogue: O / l ---------------------------------------------------------------------------------------
    mbar_wait(&pv_done[t], (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int qrow = q0 + t * 128 + r;
    const bool row_ok = qrow < p.nq;
    T* optr = reinterpret_cast<T*>(p.out) + static_cast<long long>(b) * p.bso + static_cast<long long>(qrow) * p.ldo +
              head * p.d;
#pragma unroll 1
    for (int c = 0; c < p.dN; c += 16) {
      uint32_t o[16];
      tmem_ld_x16(t_o + c, o);
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
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

template <typename T, int NA, int QT>
static int launch_attention2(const EsAttention* a, cudaStream_t stream) {
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
    const uint32_t boxkv[4] = {64u, 1u, static_cast<uint32_t>(kAtt2KV), 1u};
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
  p.o_stride = (p.dN + 31) / 32 * 32;
  const int need = QT * (kAtt2KV + p.o_stride);
  p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
  ES_CHECK(need <= 512, "es_attention: TMEM budget exceeded (d %d, %d query tiles)", a->d, QT);
  // Q + P per tile, 2 stages of K and V, barriers
  const size_t smem = static_cast<size_t>(QT) * (NA + 1) * kAtomBytes + 4 * NA * kAtt2KV * 128 + 256;
  auto kern = attention2_kernel<T, NA, QT>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    ES_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  dim3 grid((a->nq + 128 * QT - 1) / (128 * QT), a->heads, a->batch);
  ES_CUDA(launch_kernel(kern, dim3(grid), dim3(64 + 128 * QT), smem, stream, tmQ, tmK, tmV, p));
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es
