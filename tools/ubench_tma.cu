// Micro-benchmark behind the small-M GEMM design ("small-M, long-K layers"): how fast can the CTAs of a
// 512 x 1280 x 11520 GEMM (the 8x8-level 3x3 convolutions at 8 images) pull their operand tiles through TMA when
// NOTHING consumes them -- i.e. is the operand stream itself the limit, and what does TMA multicast across the four
// M tiles that share a weight tile buy?
//   mode 0  unicast, as gemm.cu does it: the 4 M-tile CTAs of an (n, split) each load the same B tile
//   mode 1  unicast, no sharing: every CTA reads its OWN copy of B (4x the DRAM bytes) -- what L2 dedup is worth
//   mode 2  cluster (4,1,1) along M: CTA r loads rows [r BN/4, +BN/4) of the B tile and multicasts them to all four
//   mode 3  as 0 without the A loads
// grid (4, 1280 / BN, splits), one producer thread and one consumer thread (waits full, arrives empty) per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tma tools/ubench_tma.cu && tools/ubench_tma
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "../edgestyle_b200/csrc/ptx.cuh"

using namespace es;

constexpr int kM = 512, kN = 1280, kK = 11520, kC = 1280;
constexpr int kABytes = 128 * 128;

struct P {
  int bn, stages, kb_total, kb_per, mode;
  long long* cyc;
};

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                      const P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = kABytes + p.bn * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  const int warp = threadIdx.x >> 5;
  const int kb0 = blockIdx.z * p.kb_per, kb1 = min(p.kb_total, kb0 + p.kb_per);
  const int m_tile = blockIdx.x, n0 = blockIdx.y * p.bn;
  const uint32_t rank = MODE == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MODE == 2 ? 4 : 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (MODE == 2) cluster_sync_all();
  const long long t0 = clock64();
  if (warp == 0) {
    if (elect_one()) {
      for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, s = it % p.stages;
        mbar_wait(&empty[s], ((it / p.stages) & 1) ^ 1);
        uint8_t* sa = smem + s * stage_bytes;
        const int bytes = (MODE == 3 ? 0 : kABytes) + p.bn * 128;
        mbar_expect_tx(&full[s], bytes);
        if (MODE != 3) tma_load_2d(sa, &tmA, &full[s], (kb % (kC / 64)) * 64, m_tile * 128);
        if (MODE == 2) {
          const int q = p.bn / 4;
          tma_load_2d_mc(sa + kABytes + rank * q * 128, &tmB, &full[s], kb * 64, n0 + rank * q, 0xf);
        } else {
          tma_load_2d(sa + kABytes, &tmB, &full[s], kb * 64, (MODE == 1 ? m_tile * kN : 0) + n0);
        }
      }
    }
  } else {
    if (elect_one()) {
      for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, s = it % p.stages;
        mbar_wait(&full[s], (it / p.stages) & 1);
        if (MODE == 2) {
          for (uint32_t r = 0; r < 4; ++r) mbar_arrive_cluster(mapa_u32(&empty[s], r));
        } else {
          mbar_arrive(&empty[s]);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && p.cyc) p.cyc[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = clock64() - t0;
  if (MODE == 2) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it / arrive on its barriers
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make2d(EncodeTiledFn fn, void* base, uint64_t inner, uint64_t outer, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {inner * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

template <int MODE>
static float run(const CUtensorMap& tmA, const CUtensorMap& tmB, P p, int splits, void* flush, size_t flush_bytes, bool cold) {
  const size_t smem = p.stages * (kABytes + p.bn * 128) + 2 * p.stages * 8 + 1024;
  cudaFuncSetAttribute(stream_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4, kN / p.bn, splits);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = MODE == 2 ? 4 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f, ts[7];
  for (int rep = 0; rep < 7; ++rep) {
    if (cold) cudaMemsetAsync(flush, rep, flush_bytes);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel<MODE>, tmA, tmB, p);
    cudaEventRecord(e1);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      exit(1);
    }
    cudaEventElapsedTime(&ts[rep], e0, e1);
    if (ts[rep] < best) best = ts[rep];
  }
  for (int i = 0; i < 7; ++i)
    for (int j = i + 1; j < 7; ++j)
      if (ts[j] < ts[i]) {
        float t = ts[i];
        ts[i] = ts[j];
        ts[j] = t;
      }
  return ts[3] * 1e3f;  // median, us
}

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fnp);
  void *a, *b, *flush;
  const size_t flush_bytes = 512ull << 20;
  cudaMalloc(&a, (size_t)kM * kC * 2);
  cudaMalloc(&b, (size_t)4 * kN * kK * 2);
  cudaMalloc(&flush, flush_bytes);
  cudaMemset(a, 0, (size_t)kM * kC * 2);
  cudaMemset(b, 0, (size_t)4 * kN * kK * 2);
  const CUtensorMap tmA = make2d(fn, a, kC, kM, 128);
  printf("M %d N %d K %d: B %.1f MB unique; us = median of 7, cold = L2 flushed before the launch\n", kM, kN, kK, kN * (double)kK * 2e-6);
  printf("%-6s %4s %6s %6s | %9s %9s | %12s %12s\n", "mode", "bn", "stages", "splits", "cold us", "warm us", "cold L2->SM", "warm L2->SM");
  const int bns[] = {128, 256};
  for (int mode = 0; mode < 4; ++mode)
    for (int bn : bns)
      for (int splits : {4, 7, 14})
        for (int stages : {3, 4, 6}) {
          if (stages * (kABytes + bn * 128) > 220 * 1024) continue;
          if (4 * (kN / bn) * splits > 296) continue;
          P p;
          p.bn = bn;
          p.stages = stages;
          p.kb_total = kK / 64;
          p.kb_per = (p.kb_total + splits - 1) / splits;
          p.mode = mode;
          p.cyc = nullptr;
          const CUtensorMap tmB = make2d(fn, b, kK, (uint64_t)(mode == 1 ? 4 : 1) * kN, mode == 2 ? bn / 4 : bn);
          float c, w;
          if (mode == 0) c = run<0>(tmA, tmB, p, splits, flush, flush_bytes, true), w = run<0>(tmA, tmB, p, splits, flush, flush_bytes, false);
          else if (mode == 1) c = run<1>(tmA, tmB, p, splits, flush, flush_bytes, true), w = run<1>(tmA, tmB, p, splits, flush, flush_bytes, false);
          else if (mode == 2) c = run<2>(tmA, tmB, p, splits, flush, flush_bytes, true), w = run<2>(tmA, tmB, p, splits, flush, flush_bytes, false);
          else c = run<3>(tmA, tmB, p, splits, flush, flush_bytes, true), w = run<3>(tmA, tmB, p, splits, flush, flush_bytes, false);
          // bytes that arrive in shared memory over all CTAs
          const double per_cta_kb = (mode == 3 ? 0 : kABytes) + bn * 128.0;
          const double tot = per_cta_kb * (kK / 64) * 4 * (kN / bn);
          printf("%-6d %4d %6d %6d | %9.1f %9.1f | %9.2f TB/s %9.2f TB/s\n", mode, bn, stages, splits, c, w, tot / c * 1e-6, tot / w * 1e-6);
          fflush(stdout);
        }
  return 0;
}
