// Per-call stages either side of the denoise loop: the two small kernels the VAE needs on
// top of es_gemm / es_groupnorm -- a row softmax for the single-head, 512-wide mid-block attention (its scores are
// produced and consumed by es_gemm; the head is wider than es_attention's 192-column TMEM budget) and the
// DiagonalGaussianDistribution sample of the encoder moments.
#include "common.h"
#include "ptx.cuh"

namespace es {

constexpr int kSmThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // `red` may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kSmThreads / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// One CTA per row: the fp32 score row is staged in shared memory once (16 B loads), exponentiated in place, and
// written back as 16-bit probabilities (8 B stores): one read and one write of the row.
template <typename T>
__global__ void __launch_bounds__(kSmThreads) softmax_rows_kernel(const float* __restrict__ s, long long lds,
                                                                  T* __restrict__ p, long long ldp, int cols,
                                                                  float scale_log2e) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: the scores are produced by the preceding GEMM
  extern __shared__ float4 row4[];
  __shared__ float red[kSmThreads / 32];
  const long long r = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(s + r * lds);
  const int n4 = cols >> 2;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n4; i += kSmThreads) {
    const float4 v = src[i];
    row4[i] = v;
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = block_reduce(m, true, red);
  const float off = m * scale_log2e;
  float sum = 0.f;
  for (int i = threadIdx.x; i < n4; i += kSmThreads) {  // each thread revisits its own elements: no barrier needed
    float4 v = row4[i];
    v.x = exp2f(v.x * scale_log2e - off);
    v.y = exp2f(v.y * scale_log2e - off);
    v.z = exp2f(v.z * scale_log2e - off);
    v.w = exp2f(v.w * scale_log2e - off);
    row4[i] = v;
    sum += (v.x + v.y) + (v.z + v.w);
  }
  sum = block_reduce(sum, false, red);
  const float inv = 1.0f / sum;
  uint2* dst = reinterpret_cast<uint2*>(p + r * ldp);
  for (int i = threadIdx.x; i < n4; i += kSmThreads) {
    const float4 v = row4[i];
    uint2 o;
    o.x = Cvt<T>::pack2(v.x * inv, v.y * inv);
    o.y = Cvt<T>::pack2(v.z * inv, v.w * inv);
    dst[i] = o;
  }
}

// z[img][ch][p] = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scale over moments [img*hw + p][ldm] whose
// columns are (mean[0..L), logvar[0..L)); noise == nullptr gives the mode.
__global__ void gaussian_sample_kernel(const float* __restrict__ moments, long long ldm, const float* __restrict__ noise,
                                       float* __restrict__ out, int n, int L, int hw, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(n) * L * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % hw);
    const int ch = static_cast<int>((i / hw) % L);
    const long long img = i / (static_cast<long long>(hw) * L);
    const float* mrow = moments + (img * hw + p) * ldm;
    float z = mrow[ch];
    if (noise) {
      const float logvar = fminf(fmaxf(mrow[L + ch], -30.f), 20.f);
      z += expf(0.5f * logvar) * noise[i];
    }
    out[i] = z * scale;
  }
}

}  // namespace es

using namespace es;

extern "C" int es_softmax_rows(int dtype, const float* s, long long lds, void* p, long long ldp, int rows, int cols,
                               float scale, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(s && p, "es_softmax_rows: null pointer");
  ES_CHECK(rows > 0 && cols > 0 && cols % 4 == 0 && lds % 4 == 0 && ldp % 4 == 0,
           "es_softmax_rows: cols and pitches must be multiples of 4 (cols=%d)", cols);
  ES_CHECK((reinterpret_cast<uintptr_t>(s) & 15) == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0,
           "es_softmax_rows: unaligned pointer");
  const size_t smem = static_cast<size_t>(cols) * sizeof(float);
  ES_CHECK(smem <= 200 * 1024, "es_softmax_rows: row of %d columns does not fit in shared memory", cols);
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == ES_DTYPE_BF16) {
    if (smem > 40 * 1024)  // 48 KB default limit covers dynamic + static shared memory
      ES_CUDA(cudaFuncSetAttribute(softmax_rows_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    ES_CUDA(launch_kernel(softmax_rows_kernel<__nv_bfloat16>, dim3(rows), dim3(kSmThreads), smem, st, s, lds,
                          reinterpret_cast<__nv_bfloat16*>(p), ldp, cols, sl2));
  } else {
    if (smem > 40 * 1024)  // 48 KB default limit covers dynamic + static shared memory
      ES_CUDA(cudaFuncSetAttribute(softmax_rows_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    ES_CUDA(launch_kernel(softmax_rows_kernel<__half>, dim3(rows), dim3(kSmThreads), smem, st, s, lds,
                          reinterpret_cast<__half*>(p), ldp, cols, sl2));
  }
  ES_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int es_gaussian_sample(const float* moments, long long ldm, const float* noise, float* out, int n,
                                  int latent_channels, int hw, float scale, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(moments && out, "es_gaussian_sample: null pointer");
  ES_CHECK(n > 0 && latent_channels > 0 && hw > 0 && ldm >= 2ll * latent_channels, "es_gaussian_sample: bad shape");
  const long long total = static_cast<long long>(n) * latent_channels * hw;
  long long g = (total + 255) / 256;
  if (g > 148ll * 16) g = 148ll * 16;
  ES_CUDA(launch_kernel(gaussian_sample_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, st, moments, ldm, noise,
                        out, n, latent_channels, hw, scale));
  ES_CUDA(cudaGetLastError());
  return 0;
}
