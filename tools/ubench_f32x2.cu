// Micro-benchmark: does the packed fp32 pipe of sm_100a (fma.rn.f32x2 -> FFMA2) raise the fp32 rate of an SM, or only
// halve the issue slots?  One CTA per SM, W warps, each thread runs 8 independent FMA chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_f32x2 tools/ubench_f32x2.cu && tools/ubench_f32x2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2(float a) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }

// MODE 0: 16 scalar FMA chains; 1: 8 packed chains (same 16 flop-lanes); 2: 16 scalar FMA + 8 ex2 per iteration; 3: 8 packed + 8 ex2
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, float* sink, long long* cycles) {
  float a[16];
  uint64_t p[8];
  float e[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]), e[i] = -0.01f * (i + threadIdx.x);
  const float s = 0.999f, t = 0.0001f;
  const uint64_t ss = pk(s, s), tt = pk(t, t);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], s, t);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, tt);
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = ex2(e[i]);
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float x, y;
    upk(p[i], x, y);
    acc += x + y + e[i];
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int warps, float* sink, long long* cyc) {
  const int iters = 4096;
  k<MODE><<<148, warps * 32>>>(iters, sink, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(iters, sink, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  const double fma_lanes = 16.0 * 32 * warps * iters;  // scalar-equivalent FMAs per SM
  printf("%-34s warps %2d: %8.0f cycles, %6.1f fp32 FMA / clk / SM%s\n", name, warps, c, fma_lanes / c,
         MODE >= 2 ? "  (+ 8 ex2 per 16 FMA)" : "");
}

int main() {
  float* sink;
  long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (int w : {4, 8, 16, 32}) {
    run<0>("scalar fma.rn.f32", w, sink, cyc);
    run<1>("packed fma.rn.f32x2", w, sink, cyc);
    run<2>("scalar fma + ex2", w, sink, cyc);
    run<3>("packed fma + ex2", w, sink, cyc);
  }
  return 0;
}
