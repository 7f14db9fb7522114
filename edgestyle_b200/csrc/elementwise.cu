// Small bandwidth / latency kernels around the tensor-core path: layout conversion at the NCHW
// boundary, im2col for the stride-2 and 4-channel convs, nearest upsample, tiny-M linears (time
// embedding), timestep sinusoid, CFG combine + DDIM update.
#include "common.h"
#include "ptx.cuh"

namespace es {

// ---------------------------------------------------------------- NCHW fp32 <-> NHWC 16-bit
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int c, int hw, long long ldd) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  // grid (ceil(hw/32), ceil(ldd/32), n); block (32, 8): smem transpose tile 32 px x 32 ch
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ch = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (ch < c && p < hw) ? src[(static_cast<long long>(n) * c + ch) * hw + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, ch = c0 + threadIdx.x;
    if (p < hw && ch < ldd) dst[(static_cast<long long>(n) * hw + p) * ldd + ch] = Cvt<T>::from_f(tile[threadIdx.x][i]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, long long lds, float* __restrict__ dst, int c, int hw) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, ch = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && ch < c) ? Cvt<T>::to_f(src[(static_cast<long long>(n) * hw + p) * lds + ch]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int ch = c0 + i, p = p0 + threadIdx.x;
    if (ch < c && p < hw) dst[(static_cast<long long>(n) * c + ch) * hw + p] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------- im2col 3x3, stride s, `pad` zero rows/columns before the image
// out[(n, yo, xo)][tap*c + ch]; one thread per (row, tap, 8-channel vector) when c % 8 == 0, scalar otherwise.
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldo, int n,
                                 int h, int w, int c, int stride, int pad, int ho, int wo) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const long long rows = static_cast<long long>(n) * ho * wo;
  if ((c & 7) == 0) {
    const int vpt = c >> 3;
    const long long total = rows * 9 * vpt;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int v = static_cast<int>(i % vpt);
      const int tap = static_cast<int>((i / vpt) % 9);
      const long long row = i / (9 * vpt);
      const int xo = static_cast<int>(row % wo), yo = static_cast<int>((row / wo) % ho);
      const int img = static_cast<int>(row / (static_cast<long long>(wo) * ho));
      const int y = yo * stride + tap / 3 - pad, x = xo * stride + tap % 3 - pad;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < h && x >= 0 && x < w)
        val = *reinterpret_cast<const uint4*>(src + ((static_cast<long long>(img) * h + y) * w + x) * lds + v * 8);
      *reinterpret_cast<uint4*>(dst + row * ldo + tap * c + v * 8) = val;
    }
  } else {
    const long long total = rows * ldo;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int col = static_cast<int>(i % ldo);
      const long long row = i / ldo;
      T val = Cvt<T>::from_f(0.f);
      if (col < 9 * c) {
        const int tap = col / c, ch = col % c;
        const int xo = static_cast<int>(row % wo), yo = static_cast<int>((row / wo) % ho);
        const int img = static_cast<int>(row / (static_cast<long long>(wo) * ho));
        const int y = yo * stride + tap / 3 - pad, x = xo * stride + tap % 3 - pad;
        if (y >= 0 && y < h && x >= 0 && x < w) val = src[((static_cast<long long>(img) * h + y) * w + x) * lds + ch];
      }
      dst[i] = val;
    }
  }
}

template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, int n,
                                  int h, int w, int c) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int vpp = c >> 3;
  const long long total = static_cast<long long>(n) * (2 * h) * (2 * w) * vpp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    const long long pix = i / vpp;
    const int xo = static_cast<int>(pix % (2 * w)), yo = static_cast<int>((pix / (2 * w)) % (2 * h));
    const int img = static_cast<int>(pix / (4ll * w * h));
    const uint4 val =
        *reinterpret_cast<const uint4*>(src + ((static_cast<long long>(img) * h + yo / 2) * w + xo / 2) * lds + v * 8);
    *reinterpret_cast<uint4*>(dst + pix * ldd + v * 8) = val;
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, long long lda, const T* __restrict__ b, long long ldb,
                           T* __restrict__ out, long long ldo, long long rows, int c) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int vpp = c >> 3;
  const long long total = rows * vpp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    const long long r = i / vpp;
    const uint4 ua = *reinterpret_cast<const uint4*>(a + r * lda + v * 8);
    const uint4 ub = *reinterpret_cast<const uint4*>(b + r * ldb + v * 8);
    const uint32_t aa[4] = {ua.x, ua.y, ua.z, ua.w}, bb[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t oo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = Cvt<T>::unpack2(aa[j]), fb = Cvt<T>::unpack2(bb[j]);
      oo[j] = Cvt<T>::pack2(fa.x + fb.x, fa.y + fb.y);
    }
    *reinterpret_cast<uint4*>(out + r * ldo + v * 8) = make_uint4(oo[0], oo[1], oo[2], oo[3]);
  }
}

// ---------------------------------------------------------------- timestep sinusoid
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int n, int dim, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, k = i % half;
  const float freq = expf(-logf(10000.0f) * static_cast<float>(k) / static_cast<float>(half));
  const float a = t[r] * freq;
  out[r * dim + k] = cosf(a);          // flip_sin_to_cos: [cos | sin]
  out[r * dim + half + k] = sinf(a);
}

// ---------------------------------------------------------------- tiny-M linear: one warp per two output columns
// x rows (activation applied once) are staged in shared memory; every lane streams 16 B weight vectors of its two
// columns, so a block keeps 2 x 8 independent weight streams in flight (HBM-bound GEMV: the weights are read once).
constexpr int kSlWarps = 8;
constexpr int kSlCols = 2;
template <typename T, int RC>
__global__ void small_linear_kernel(const float* __restrict__ x, int ldx, const T* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y, int ldy, int rows, int n,
                                    int k, int silu_in, int silu_out, int accumulate) {
  extern __shared__ float xs[];  // [min(rows, RC)][k]
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int lane = threadIdx.x & 31;
  const int col0 = (blockIdx.x * kSlWarps + (threadIdx.x >> 5)) * kSlCols;
  for (int r0 = 0; r0 < rows; r0 += RC) {
    const int nr = min(RC, rows - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * k; i += blockDim.x) {
      const int r = i / k, c = i - r * k;
      float xv = x[static_cast<long long>(r0 + r) * ldx + c];
      xs[i] = silu_in ? silu_f(xv) : xv;
    }
    __syncthreads();
    float acc[kSlCols][RC];
#pragma unroll
    for (int cc = 0; cc < kSlCols; ++cc)
#pragma unroll
      for (int r = 0; r < RC; ++r) acc[cc][r] = 0.f;
    for (int kk = lane * 8; kk < k; kk += 256) {
      uint4 u[kSlCols];
#pragma unroll
      for (int cc = 0; cc < kSlCols; ++cc)
        u[cc] = (col0 + cc < n) ? *reinterpret_cast<const uint4*>(w + static_cast<long long>(col0 + cc) * k + kk)
                                : make_uint4(0, 0, 0, 0);
      float wf[kSlCols][8];
#pragma unroll
      for (int cc = 0; cc < kSlCols; ++cc) {
        const uint32_t uu[4] = {u[cc].x, u[cc].y, u[cc].z, u[cc].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = Cvt<T>::unpack2(uu[j]);
          wf[cc][2 * j] = f.x;
          wf[cc][2 * j + 1] = f.y;
        }
      }
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        if (r < nr) {
          const float4 x0 = *reinterpret_cast<const float4*>(xs + r * k + kk);
          const float4 x1 = *reinterpret_cast<const float4*>(xs + r * k + kk + 4);
          const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int cc = 0; cc < kSlCols; ++cc)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[cc][r] += xv[j] * wf[cc][j];
        }
      }
    }
#pragma unroll
    for (int cc = 0; cc < kSlCols; ++cc)
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        float a = acc[cc][r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && r < nr && col0 + cc < n) {
          if (bias) a += bias[col0 + cc];
          if (silu_out) a = silu_f(a);
          float* yp = y + static_cast<long long>(r0 + r) * ldy + col0 + cc;
          *yp = accumulate ? (*yp + a) : a;
        }
      }
  }
}

// ---------------------------------------------------------------- CFG + DDIM
__global__ void cfg_ddim_kernel(const float* __restrict__ eps, float* __restrict__ lat,
                                const float* __restrict__ guidance, const float* __restrict__ coef,
                                float* __restrict__ eps_out, int imgs, int chw) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(imgs) * chw) return;
  const int img = static_cast<int>(i / chw);
  const float eu = eps[i];
  const float ec = eps[i + static_cast<long long>(imgs) * chw];
  const float e = eu + guidance[img] * (ec - eu);
  if (eps_out) eps_out[i] = e;
  const float sa = coef[0], s1a = coef[1], sp = coef[2], s1p = coef[3];
  const float x = lat[i];
  const float x0 = (x - s1a * e) / sa;
  lat[i] = sp * x0 + s1p * e;
}

// x0-prediction from the CFG-combined noise prediction (UniPC `convert_model_output`, predict_x0):
// eps = eu + g (ec - eu);  x0 = (sample - sigma * eps) / alpha
__global__ void cfg_x0_kernel(const float* __restrict__ eps, const float* __restrict__ sample,
                              const float* __restrict__ guidance, float alpha, float sigma, float* __restrict__ x0,
                              int imgs, int chw) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(imgs) * chw) return;
  const int img = static_cast<int>(i / chw);
  const float eu = eps[i];
  const float ec = eps[i + static_cast<long long>(imgs) * chw];
  const float e = eu + guidance[img] * (ec - eu);
  x0[i] = (sample[i] - sigma * e) / alpha;
}

// out = c0*x0 + c1*x1 + c2*x2 + c3*x3 (null pointers skipped): the UniPC predictor / corrector updates are linear
// combinations of {sample, last_sample, model outputs} with host-computed scalar coefficients.
__global__ void lincomb4_kernel(float* __restrict__ out, float c0, const float* __restrict__ x0, float c1,
                                const float* __restrict__ x1, float c2, const float* __restrict__ x2, float c3,
                                const float* __restrict__ x3, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float v = 0.f;
  if (x0) v += c0 * x0[i];
  if (x1) v += c1 * x1[i];
  if (x2) v += c2 * x2[i];
  if (x3) v += c3 * x3[i];
  out[i] = v;
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148ll * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace es

using namespace es;

extern "C" int es_nchw_to_nhwc(int dtype, const float* src, void* dst, int n, int c, int hw, long long ldd,
                               void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ceil_div(hw, 32), ceil_div(static_cast<int>(ldd), 32), n), block(32, 8);
  if (dtype == ES_DTYPE_BF16)
    ES_CUDA(launch_kernel(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, s, src, reinterpret_cast<__nv_bfloat16*>(dst), c, hw, ldd));
  else
    ES_CUDA(launch_kernel(nchw_to_nhwc_kernel<__half>, dim3(grid), dim3(block), 0, s, src, reinterpret_cast<__half*>(dst), c, hw, ldd));
  ES_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int es_nhwc_to_nchw(int dtype, const void* src, long long lds, float* dst, int n, int c, int hw,
                               void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ceil_div(hw, 32), ceil_div(c, 32), n), block(32, 8);
  if (dtype == ES_DTYPE_BF16)
    ES_CUDA(launch_kernel(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, s, reinterpret_cast<const __nv_bfloat16*>(src), lds, dst, c, hw));
  else
    ES_CUDA(launch_kernel(nhwc_to_nchw_kernel<__half>, dim3(grid), dim3(block), 0, s, reinterpret_cast<const __half*>(src), lds, dst, c, hw));
  ES_CUDA(cudaGetLastError());
  return 0;
}
static int im2col3x3_launch(int dtype, const void* src, long long lds, void* dst, long long ldo, int n, int h, int w,
                            int c, int stride, int pad_lo, int pad_hi, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(stride == 1 || stride == 2, "es_im2col3x3: stride must be 1 or 2");
  ES_CHECK(pad_lo >= 0 && pad_lo <= 1 && pad_hi >= 0 && pad_hi <= 1, "es_im2col3x3: padding must be 0 or 1 per side");
  ES_CHECK(h + pad_lo + pad_hi >= 3 && w + pad_lo + pad_hi >= 3, "es_im2col3x3: image smaller than the window");
  ES_CHECK(ldo >= 9ll * c, "es_im2col3x3: ldo too small");
  if (c % 8 == 0) ES_CHECK(lds % 8 == 0 && ldo == 9ll * c, "es_im2col3x3: vector path needs lds%%8==0 and ldo==9c");
  const int ho = (h + pad_lo + pad_hi - 3) / stride + 1, wo = (w + pad_lo + pad_hi - 3) / stride + 1;
  const long long total = static_cast<long long>(n) * ho * wo * ((c % 8 == 0) ? 9 * (c / 8) : ldo);
  const int g = grid_for(total, 256);
  if (dtype == ES_DTYPE_BF16)
    ES_CUDA(launch_kernel(im2col3x3_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(src), lds,
                                                      reinterpret_cast<__nv_bfloat16*>(dst), ldo, n, h, w, c, stride, pad_lo, ho, wo));
  else
    ES_CUDA(launch_kernel(im2col3x3_kernel<__half>, dim3(g), dim3(256), 0, s, reinterpret_cast<const __half*>(src), lds, reinterpret_cast<__half*>(dst),
                                               ldo, n, h, w, c, stride, pad_lo, ho, wo));
  ES_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int es_im2col3x3(int dtype, const void* src, long long lds, void* dst, long long ldo, int n, int h, int w,
                            int c, int stride, void* stream) {
  return im2col3x3_launch(dtype, src, lds, dst, ldo, n, h, w, c, stride, 1, 1, stream);
}
extern "C" int es_im2col3x3_pad(int dtype, const void* src, long long lds, void* dst, long long ldo, int n, int h,
                                int w, int c, int stride, int pad_lo, int pad_hi, void* stream) {
  return im2col3x3_launch(dtype, src, lds, dst, ldo, n, h, w, c, stride, pad_lo, pad_hi, stream);
}
extern "C" int es_upsample2x(int dtype, const void* src, long long lds, void* dst, long long ldd, int n, int h, int w,
                             int c, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(c % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "es_upsample2x: c and pitches must be multiples of 8");
  const long long total = static_cast<long long>(n) * 4 * h * w * (c / 8);
  const int g = grid_for(total, 256);
  if (dtype == ES_DTYPE_BF16)
    ES_CUDA(launch_kernel(upsample2x_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(src), lds,
                                                       reinterpret_cast<__nv_bfloat16*>(dst), ldd, n, h, w, c));
  else
    ES_CUDA(launch_kernel(upsample2x_kernel<__half>, dim3(g), dim3(256), 0, s, reinterpret_cast<const __half*>(src), lds, reinterpret_cast<__half*>(dst),
                                                ldd, n, h, w, c));
  ES_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int es_add(int dtype, const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo,
                      int rows, int c, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(c % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0, "es_add: c and pitches must be multiples of 8");
  const long long total = static_cast<long long>(rows) * (c / 8);
  const int g = grid_for(total, 256);
  if (dtype == ES_DTYPE_BF16)
    ES_CUDA(launch_kernel(add_kernel<__nv_bfloat16>, dim3(g), dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(a), lda,
                                                reinterpret_cast<const __nv_bfloat16*>(b), ldb,
                                                reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, c));
  else
    ES_CUDA(launch_kernel(add_kernel<__half>, dim3(g), dim3(256), 0, s, reinterpret_cast<const __half*>(a), lda, reinterpret_cast<const __half*>(b), ldb,
                                         reinterpret_cast<__half*>(out), ldo, rows, c));
  ES_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int es_timestep_embedding(const float* t, int n, int dim, float* out, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(dim % 2 == 0, "es_timestep_embedding: dim must be even");
  const int total = n * dim / 2;
  ES_CUDA(launch_kernel(timestep_embedding_kernel, dim3(ceil_div(total, 128)), dim3(128), 0, s, t, n, dim, out));
  ES_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int es_small_linear(int dtype, const float* x, int ldx, const void* w, const float* bias, float* y, int ldy,
                               int rows, int n, int k, int silu_in, int silu_out, int accumulate, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CHECK(k % 8 == 0 && ldx % 4 == 0, "es_small_linear: k must be a multiple of 8");
  constexpr int RC = 4;
  const size_t smem = static_cast<size_t>(rows < RC ? rows : RC) * k * sizeof(float);
  ES_CHECK(smem <= 48 * 1024, "es_small_linear: k too large (%d)", k);
  dim3 grid(ceil_div(n, kSlWarps * kSlCols)), block(kSlWarps * 32);
  if (dtype == ES_DTYPE_BF16)
    ES_CUDA(launch_kernel(small_linear_kernel<__nv_bfloat16, RC>, dim3(grid), dim3(block), smem, s, x, ldx,
                          reinterpret_cast<const __nv_bfloat16*>(w), bias, y, ldy, rows, n, k, silu_in, silu_out, accumulate));
  else
    ES_CUDA(launch_kernel(small_linear_kernel<__half, RC>, dim3(grid), dim3(block), smem, s, x, ldx,
                          reinterpret_cast<const __half*>(w), bias, y, ldy, rows, n, k, silu_in, silu_out, accumulate));
  ES_CUDA(cudaGetLastError());
  return 0;
}
// Timeline probe: writes %globaltimer (ns) after every earlier kernel of the stream has completed.
__global__ void stamp_kernel(unsigned long long* slot) {
  pdl_wait();
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}
extern "C" int es_stamp(unsigned long long* slot, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CUDA(launch_kernel(stamp_kernel, dim3(1), dim3(1), 0, s, slot));
  return 0;
}
extern "C" int es_cfg_x0(const float* eps, const float* sample, const float* guidance, float alpha, float sigma, float* x0,
                         int imgs, int chw, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(imgs) * chw;
  ES_CUDA(launch_kernel(cfg_x0_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, eps, sample, guidance,
                        alpha, sigma, x0, imgs, chw));
  return 0;
}
extern "C" int es_lincomb4(float* out, float c0, const float* x0, float c1, const float* x1, float c2, const float* x2,
                           float c3, const float* x3, long long n, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ES_CUDA(launch_kernel(lincomb4_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, s, out, c0, x0, c1, x1, c2,
                        x2, c3, x3, n));
  return 0;
}
extern "C" int es_cfg_ddim(const float* eps, float* latents, const float* guidance, const float* coef, float* eps_out,
                           int imgs, int chw, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(imgs) * chw;
  ES_CUDA(launch_kernel(cfg_ddim_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, eps, latents, guidance, coef, eps_out, imgs, chw));
  ES_CUDA(cudaGetLastError());
  return 0;
}
