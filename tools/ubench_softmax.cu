// Micro-benchmark behind the attention design (kernels: es_attention): per-SM throughput of the three
// things a flash-attention softmax warp does with one 128 x 64 score tile on sm_100a --
//   (a) ex2.approx.ftz.f32   (b) ex2.approx.f16x2   (c) tcgen05.ld 32x32b.x32 (TMEM -> registers)
// alone and together, to see which of them add up.  One CTA per SM, `warps` warps each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_softmax tools/ubench_softmax.cu && tools/ubench_softmax
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float* sink, long long* cycles) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float f[32];
  uint32_t h[32];
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    f[i] = -0.001f * (threadIdx.x + i);
    h[i] = 0xb800b400u + i;  // two small negative halves
    v[i] = 0;
  }
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2 || MODE == 3 || MODE == 4 || (MODE == 5 && (warp & 4))) {  // TMEM load of 32 columns (4 KB per warp)
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr + (it & 1) * 32)
          : "memory");
    }
    if (MODE == 0 || MODE == 3 || (MODE == 5 && !(warp & 4))) {  // 32 fp32 exponentials
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
    }
    if (MODE == 1 || MODE == 4) {  // 16 packed exponentials = 32 values
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
    if (MODE == 2 || MODE == 3 || MODE == 4 || (MODE == 5 && (warp & 4))) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(v[0] & 0x3fffffffu) + __uint_as_float(v[31] & 0x3fffffffu);
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += f[i] + __uint_as_float(h[i] & 0x3fffffffu);
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(128));
}

// TMEM load shapes: bytes per warp-instruction and cycles, 1 and 4 warps per sub-partition
template <int SHAPE>  // 0: 32x32b.x16  1: 32x32b.x64  2: 16x256b.x4 (16 regs)  3: 16x256b.x8 (32 regs)  4: 32x32b.x8  5: 16x128b.x8 (16 regs)
__global__ void __launch_bounds__(512, 1) kld(int iters, float* sink, long long* cycles) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (SHAPE == 0) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr) : "memory");
    } else if (SHAPE == 4) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr) : "memory");
    } else if (SHAPE == 2) {
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr) : "memory");
    } else if (SHAPE == 5) {
      asm volatile("tcgen05.ld.sync.aligned.16x128b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr) : "memory");
    } else if (SHAPE == 1) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
          "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
            "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
            "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
            "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
          : "r"(taddr) : "memory");
    } else {
      asm volatile(
          "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr) : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(v[0] & 0x3fffffffu) + __uint_as_float(v[7] & 0x3fffffffu);
  }
  const long long t1 = clock64();
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(128));
}

template <int SHAPE>
static void run_ld(const char* name, int bytes_per_warp, int warps, int iters, float* sink, long long* cyc_d) {
  kld<SHAPE><<<148, warps * 32>>>(iters, sink, cyc_d);
  cudaDeviceSynchronize();
  kld<SHAPE><<<148, warps * 32>>>(iters, sink, cyc_d);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc[148];
  cudaMemcpy(cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += cyc[i];
  avg /= 148;
  const double per = avg / iters / (warps / 4.0);
  printf("%-28s warps/SM %2d  %7.1f cycles per load+wait  %6.1f B/clk/SM  %s\n", name, warps, per, 4.0 * bytes_per_warp / per,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int MODE>
static void run(const char* name, int warps, int iters, float* sink, long long* cyc_d) {
  k<MODE><<<148, warps * 32>>>(iters, sink, cyc_d);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(iters, sink, cyc_d);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc[148];
  cudaMemcpy(cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += cyc[i];
  avg /= 148;
  // per SM sub-partition: warps / 4 warps share it; one iteration = 32 scores per lane
  printf("%-44s warps/SM %2d  cycles/iter/SMSP-warp-slot %8.1f  (per warp-iteration: %7.1f)  %s\n", name, warps,
         avg / iters, avg / iters / (warps / 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 16);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    run<0>("(a) 32 x ex2.f32", warps, iters, sink, cyc);
    run<1>("(b) 16 x ex2.f16x2 (32 values)", warps, iters, sink, cyc);
    run<2>("(c) tcgen05.ld 32x32b.x32", warps, iters, sink, cyc);
    run<3>("(a)+(c) same warp", warps, iters, sink, cyc);
    run<4>("(b)+(c) same warp", warps, iters, sink, cyc);
    if (warps >= 8) run<5>("(a) | (c) on different warps of an SMSP", warps, iters, sink, cyc);
  }
  for (int warps : {4, 16}) {
    run_ld<4>("tcgen05.ld 32x32b.x8", 1024, warps, iters, sink, cyc);
    run_ld<0>("tcgen05.ld 32x32b.x16", 2048, warps, iters, sink, cyc);
    run_ld<1>("tcgen05.ld 32x32b.x64", 8192, warps, iters, sink, cyc);
    run_ld<5>("tcgen05.ld 16x128b.x8", 2048, warps, iters, sink, cyc);
    run_ld<2>("tcgen05.ld 16x256b.x4", 2048, warps, iters, sink, cyc);
    run_ld<3>("tcgen05.ld 16x256b.x8", 4096, warps, iters, sink, cyc);
  }
  return 0;
}
