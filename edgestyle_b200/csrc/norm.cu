// Bandwidth kernels: GroupNorm(+SiLU)(+channel concat) and LayerNorm over channels-last activations.
// fp16/bf16 in/out, fp32 statistics, 16-byte vector loads/stores, fixed channel-vector per thread so the
// per-channel scale/shift lives in registers.
#include "common.h"
#include "ptx.cuh"

namespace es {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = Cvt<T>::unpack2(u.x), b = Cvt<T>::unpack2(u.y), c = Cvt<T>::unpack2(u.z), d = Cvt<T>::unpack2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&f)[8]) {
  uint4 u;
  u.x = Cvt<T>::pack2(f[0], f[1]); u.y = Cvt<T>::pack2(f[2], f[3]);
  u.z = Cvt<T>::pack2(f[4], f[5]); u.w = Cvt<T>::pack2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// grid (chunks, n_img); block (C/8, ny)
template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x0, int c0, long long ld0, const T* __restrict__ x1, int c1,
                                long long ld1, int hw, int groups, float* __restrict__ ws) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  extern __shared__ float sm[];  // [C][2]
  const int C = c0 + c1;
  const int cpg = C / groups;
  const int v = threadIdx.x;  // channel vector
  const int img = blockIdx.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 2 * C; i += blockDim.x * blockDim.y) sm[i] = 0.f;
  __syncthreads();
  const T* src;
  long long ld;
  int ch = v * 8;
  if (ch < c0) {
    src = x0 + static_cast<long long>(img) * hw * ld0 + ch;
    ld = ld0;
  } else {
    src = x1 + static_cast<long long>(img) * hw * ld1 + (ch - c0);
    ld = ld1;
  }
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  const int per = (hw + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per;
  const int p1 = min(hw, p0 + per);
  {
    const int ny = blockDim.y;
    int p = p0 + threadIdx.y;
    for (; p + 3 * ny < p1; p += 4 * ny) {  // four independent 16 B loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(src + static_cast<long long>(p + q * ny) * ld);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = Cvt<T>::unpack2(w[j]);
          s[2 * j] += f.x; ss[2 * j] += f.x * f.x;
          s[2 * j + 1] += f.y; ss[2 * j + 1] += f.y * f.y;
        }
      }
    }
    for (; p < p1; p += ny) {
      float f[8];
      load8<T>(src + static_cast<long long>(p) * ld, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        ss[j] += f[j] * f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sm[2 * (ch + j)], s[j]);
    atomicAdd(&sm[2 * (ch + j) + 1], ss[j]);
  }
  __syncthreads();
  for (int g = tid; g < groups; g += blockDim.x * blockDim.y) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < cpg; ++j) {
      a += sm[2 * (g * cpg + j)];
      b += sm[2 * (g * cpg + j) + 1];
    }
    atomicAdd(&ws[(static_cast<long long>(img) * groups + g) * 2], a);
    atomicAdd(&ws[(static_cast<long long>(img) * groups + g) * 2 + 1], b);
  }
}

template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x0, int c0, long long ld0, const T* __restrict__ x1, int c1,
                                long long ld1, int hw, int groups, float eps, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ ws, T* __restrict__ out,
                                long long ldo, int silu) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int C = c0 + c1;
  const int cpg = C / groups;
  const int v = threadIdx.x;
  const int img = blockIdx.y;
  const int ch = v * 8;
  const T* src;
  long long ld;
  if (ch < c0) {
    src = x0 + static_cast<long long>(img) * hw * ld0 + ch;
    ld = ld0;
  } else {
    src = x1 + static_cast<long long>(img) * hw * ld1 + (ch - c0);
    ld = ld1;
  }
  float a[8], b[8];
  const float inv_n = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (ch + j) / cpg;
    const float sum = ws[(static_cast<long long>(img) * groups + g) * 2];
    const float sq = ws[(static_cast<long long>(img) * groups + g) * 2 + 1];
    const float mean = sum * inv_n;
    const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    a[j] = rstd * gamma[ch + j];
    b[j] = beta[ch + j] - mean * a[j];
  }
  T* dst = out + static_cast<long long>(img) * hw * ldo + ch;
  const int per = (hw + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per;
  const int p1 = min(hw, p0 + per);
  const int ny = blockDim.y;
  int p = p0 + threadIdx.y;
  for (; p + 3 * ny < p1; p += 4 * ny) {
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(src + static_cast<long long>(p + q * ny) * ld);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = Cvt<T>::unpack2(w[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = f[j] * a[j] + b[j];
        f[j] = silu ? silu_f(y) : y;
      }
      store8<T>(dst + static_cast<long long>(p + q * ny) * ldo, f);
    }
  }
  for (; p < p1; p += ny) {
    float f[8];
    load8<T>(src + static_cast<long long>(p) * ld, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = f[j] * a[j] + b[j];
      f[j] = silu ? silu_f(y) : y;
    }
    store8<T>(dst + static_cast<long long>(p) * ldo, f);
  }
}

static int gn_check(const EsGroupNorm* g) {
  ES_CHECK(g && g->x0 && g->ws, "es_groupnorm: null pointer");
  const int C = g->c0 + g->c1;
  ES_CHECK(g->c0 % 8 == 0 && g->c1 % 8 == 0 && g->ld0 % 8 == 0 && (g->c1 == 0 || g->ld1 % 8 == 0),
           "es_groupnorm: channels and pitches must be multiples of 8");
  ES_CHECK(C % g->groups == 0 && C / 8 <= 1024, "es_groupnorm: bad channel count %d", C);
  ES_CHECK(g->c1 == 0 || g->x1, "es_groupnorm: c1 > 0 but x1 is null");
  return 0;
}

static void gn_geometry(const EsGroupNorm* g, dim3& grid, dim3& block) {
  const int vpp = (g->c0 + g->c1) / 8;
  int ny = 256 / vpp;
  if (ny < 1) ny = 1;
  if (ny > g->hw) ny = g->hw;
  block = dim3(vpp, ny, 1);
  // >= 4 pixels per thread (four independent 16 B loads in flight, per-thread prologue amortised), ~4 blocks per SM
  int chunks = (4 * 148 + g->n_img - 1) / g->n_img;
  const int max_chunks = (g->hw + ny * 4 - 1) / (ny * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  grid = dim3(chunks, g->n_img, 1);
}

template <typename T>
static int gn_stats_t(const EsGroupNorm* g, cudaStream_t s) {
  dim3 grid, block;
  gn_geometry(g, grid, block);
  const size_t smem = static_cast<size_t>(g->c0 + g->c1) * 2 * sizeof(float);
  ES_CUDA(launch_kernel(gn_stats_kernel<T>, dim3(grid), dim3(block), smem, s, reinterpret_cast<const T*>(g->x0), g->c0, g->ld0,
                                               reinterpret_cast<const T*>(g->x1), g->c1, g->ld1, g->hw, g->groups,
                                               g->ws));
  ES_CUDA(cudaGetLastError());
  return 0;
}
template <typename T>
static int gn_apply_t(const EsGroupNorm* g, cudaStream_t s) {
  dim3 grid, block;
  gn_geometry(g, grid, block);
  ES_CHECK(g->out && g->ldo % 8 == 0 && g->gamma && g->beta, "es_groupnorm_apply: bad output/affine");
  ES_CUDA(launch_kernel(gn_apply_kernel<T>, dim3(grid), dim3(block), 0, s, reinterpret_cast<const T*>(g->x0), g->c0, g->ld0,
                                            reinterpret_cast<const T*>(g->x1), g->c1, g->ld1, g->hw, g->groups, g->eps,
                                            g->gamma, g->beta, g->ws, reinterpret_cast<T*>(g->out), g->ldo, g->silu));
  ES_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- GroupNorm in ONE pass over memory (cluster + DSMEM)
// One thread-block cluster of 8 CTAs per (image, group): every CTA keeps its 1/8 of the group's hw x cpg slab in
// registers, the eight partial (sum, sumsq) meet through distributed shared memory, and the normalised (+SiLU) values
// are written from registers -- one read and one write of the tensor, one launch, no statistics buffer.  Used where
// the producer could not accumulate the statistics itself (the decoder's [x | skip] concat inputs, the first resnet
// of an encoder pass).
constexpr int kGnfThreads = 512;
constexpr int kGnfCluster = 8;

template <typename T, int PPT, int JPT>
__global__ void __cluster_dims__(kGnfCluster, 1, 1) __launch_bounds__(kGnfThreads, 1)
gn_fused_kernel(const T* __restrict__ x, int C, long long ldx, int hw, int groups, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ out, long long ldo,
                int silu) {
  __shared__ float warp_part[kGnfThreads / 32][2];
  __shared__ float cta_part[2];
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int rank = static_cast<int>(cluster_ctarank());
  const int g = blockIdx.x / kGnfCluster;
  const int img = blockIdx.y;
  const int cpg = C / groups;
  const int pairs = cpg >> 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int px_per_cta = hw / kGnfCluster;
  const int p_begin = rank * px_per_cta;
  const int ch0 = g * cpg;
  const T* src = x + (static_cast<long long>(img) * hw + p_begin) * ldx + ch0;
  uint32_t v[PPT][JPT];
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int pi = 0; pi < PPT; ++pi) {
    const int p = ty + 16 * pi;
#pragma unroll
    for (int jj = 0; jj < JPT; ++jj) {
      const int j = tx + 32 * jj;
      v[pi][jj] = 0u;
      if (p < px_per_cta && j < pairs) v[pi][jj] = *reinterpret_cast<const uint32_t*>(src + static_cast<long long>(p) * ldx + 2 * j);
    }
  }
#pragma unroll
  for (int pi = 0; pi < PPT; ++pi)
#pragma unroll
    for (int jj = 0; jj < JPT; ++jj) {
      const float2 f = Cvt<T>::unpack2(v[pi][jj]);  // elements that were not loaded are zero
      s += f.x + f.y;
      q += f.x * f.x + f.y * f.y;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (tx == 0) {
    warp_part[ty][0] = s;
    warp_part[ty][1] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < kGnfThreads / 32; ++w) {
      a += warp_part[w][0];
      b += warp_part[w][1];
    }
    cta_part[0] = a;
    cta_part[1] = b;
  }
  cluster_sync_all();  // every CTA's partial is visible cluster-wide
  float ts = 0.f, tq = 0.f;
#pragma unroll
  for (int r = 0; r < kGnfCluster; ++r) {
    const uint32_t a = mapa_u32(cta_part, static_cast<uint32_t>(r));
    float ps, pq;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(ps) : "r"(a));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(pq) : "r"(a + 4));
    ts += ps;
    tq += pq;
  }
  cluster_sync_all();  // nobody exits (and releases its shared memory) while a peer still reads it
  const float inv_n = 1.0f / (static_cast<float>(cpg) * static_cast<float>(hw));
  const float mean = ts * inv_n;
  const float rstd = rsqrtf(fmaxf(tq * inv_n - mean * mean, 0.f) + eps);
  T* dst = out + (static_cast<long long>(img) * hw + p_begin) * ldo + ch0;
#pragma unroll
  for (int jj = 0; jj < JPT; ++jj) {
    const int j = tx + 32 * jj;
    if (j >= pairs) continue;
    const float2 gm = *reinterpret_cast<const float2*>(gamma + ch0 + 2 * j);
    const float2 bt = *reinterpret_cast<const float2*>(beta + ch0 + 2 * j);
    const float a0 = rstd * gm.x, a1 = rstd * gm.y;
    const float b0 = bt.x - mean * a0, b1 = bt.y - mean * a1;
#pragma unroll
    for (int pi = 0; pi < PPT; ++pi) {
      const int p = ty + 16 * pi;
      if (p < px_per_cta) {
        const float2 f = Cvt<T>::unpack2(v[pi][jj]);
        float y0 = f.x * a0 + b0, y1 = f.y * a1 + b1;
        if (silu) {
          y0 = silu_f(y0);
          y1 = silu_f(y1);
        }
        *reinterpret_cast<uint32_t*>(dst + static_cast<long long>(p) * ldo + 2 * j) = Cvt<T>::pack2(y0, y1);
      }
    }
  }
}

template <typename T>
static int gn_fused_t(const EsGroupNorm* g, cudaStream_t s) {
  const int C = g->c0;
  const int cpg = C / g->groups;
  ES_CHECK(g->c1 == 0 && g->hw % kGnfCluster == 0 && cpg % 2 == 0 && C % 2 == 0 && g->ld0 % 2 == 0 && g->ldo % 2 == 0,
           "es_groupnorm_fused: unsupported geometry (hw %d, channels per group %d)", g->hw, cpg);
  ES_CHECK(g->out && g->gamma && g->beta, "es_groupnorm_fused: bad output/affine");
  const int ppt = ceil_div(g->hw / kGnfCluster, 16);
  const int jpt = ceil_div(cpg / 2, 32);
  dim3 grid(g->groups * kGnfCluster, g->n_img, 1), block(kGnfThreads);
  const T* x = reinterpret_cast<const T*>(g->x0);
  T* o = reinterpret_cast<T*>(g->out);
#define ES_GNF(P, J)                                                                                                  \
  ES_CUDA(launch_kernel(gn_fused_kernel<T, P, J>, dim3(grid), dim3(block), 0, s, x, C, g->ld0, g->hw, g->groups, g->eps, \
                        g->gamma, g->beta, o, g->ldo, g->silu))
  if (jpt == 1 && ppt <= 4) ES_GNF(4, 1);
  else if (jpt == 1 && ppt <= 8) ES_GNF(8, 1);
  else if (jpt == 1 && ppt <= 32) ES_GNF(32, 1);
  else if (jpt == 2 && ppt <= 2) ES_GNF(2, 2);
  else if (jpt == 2 && ppt <= 8) ES_GNF(8, 2);
  else if (jpt == 2 && ppt <= 32) ES_GNF(32, 2);
  else ES_CHECK(false, "es_groupnorm_fused: slab too large (hw %d, channels per group %d)", g->hw, cpg);
#undef ES_GNF
  ES_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- LayerNorm: warps own rows, R rows in flight
template <typename T, int MAXV, int R>
__global__ void layernorm_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ out, long long ldo,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int c,
                                 float eps) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int nv = c >> 3;
  uint4 u[R][MAXV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const T* src = x + static_cast<long long>(min(row0 + r, rows - 1)) * ldx;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int v = lane + i * 32;
      u[r][i] = v < nv ? *reinterpret_cast<const uint4*>(src + v * 8) : make_uint4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float f[MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const uint32_t w[4] = {u[r][i].x, u[r][i].y, u[r][i].z, u[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = Cvt<T>::unpack2(w[j]);
        f[i][2 * j] = t.x;
        f[i][2 * j + 1] = t.y;
        s += t.x + t.y;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[i][j] - mean;
          q += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / c + eps);
    if (row0 + r < rows) {
      T* dst = out + static_cast<long long>(row0 + r) * ldo;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + i * 32;
        if (v < nv) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * __ldg(gamma + v * 8 + j) + __ldg(beta + v * 8 + j);
          store8<T>(dst + v * 8, o);
        }
      }
    }
  }
}

template <typename T>
static int layernorm_t(const void* x, long long ldx, void* out, long long ldo, const float* gamma, const float* beta,
                       int rows, int c, float eps, cudaStream_t s) {
  ES_CHECK(c % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && c <= 8 * 32 * 8, "es_layernorm: unsupported c=%d", c);
  const int warps = 8;
  const int nv = c / 8;
  const T* xp = reinterpret_cast<const T*>(x);
  T* op = reinterpret_cast<T*>(out);
  // rows in flight per warp: more when there are plenty of rows (64x64 / 32x32 levels), else 1
  const bool many = rows >= 8192;
  const int R = many ? (nv <= 64 ? 4 : 2) : 1;
  dim3 grid((rows + warps * R - 1) / (warps * R)), block(warps * 32);
  if (nv <= 32 * 2) {
    if (many) ES_CUDA(launch_kernel(layernorm_kernel<T, 2, 4>, grid, block, 0, s, xp, ldx, op, ldo, gamma, beta, rows, c, eps));
    else ES_CUDA(launch_kernel(layernorm_kernel<T, 2, 1>, grid, block, 0, s, xp, ldx, op, ldo, gamma, beta, rows, c, eps));
  } else if (nv <= 32 * 5) {
    if (many) ES_CUDA(launch_kernel(layernorm_kernel<T, 5, 2>, grid, block, 0, s, xp, ldx, op, ldo, gamma, beta, rows, c, eps));
    else ES_CUDA(launch_kernel(layernorm_kernel<T, 5, 1>, grid, block, 0, s, xp, ldx, op, ldo, gamma, beta, rows, c, eps));
  } else {
    ES_CUDA(launch_kernel(layernorm_kernel<T, 8, 1>, grid, block, 0, s, xp, ldx, op, ldo, gamma, beta, rows, c, eps));
  }
  return 0;
}

}  // namespace es

extern "C" int es_groupnorm_stats(const EsGroupNorm* g, void* stream) {
  if (es::gn_check(g)) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return g->dtype == ES_DTYPE_BF16 ? es::gn_stats_t<__nv_bfloat16>(g, s) : es::gn_stats_t<__half>(g, s);
}
extern "C" int es_groupnorm_apply(const EsGroupNorm* g, void* stream) {
  if (es::gn_check(g)) return -1;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return g->dtype == ES_DTYPE_BF16 ? es::gn_apply_t<__nv_bfloat16>(g, s) : es::gn_apply_t<__half>(g, s);
}
extern "C" int es_groupnorm_fused(const EsGroupNorm* g, void* stream) {
  ES_CHECK(g && g->x0, "es_groupnorm_fused: null pointer");
  ES_CHECK(g->groups > 0 && g->c0 % g->groups == 0, "es_groupnorm_fused: bad channel count %d", g->c0);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return g->dtype == ES_DTYPE_BF16 ? es::gn_fused_t<__nv_bfloat16>(g, s) : es::gn_fused_t<__half>(g, s);
}
extern "C" int es_layernorm(int dtype, const void* x, long long ldx, void* out, long long ldo, const float* gamma,
                            const float* beta, int rows, int c, float eps, void* stream) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return dtype == ES_DTYPE_BF16 ? es::layernorm_t<__nv_bfloat16>(x, ldx, out, ldo, gamma, beta, rows, c, eps, s)
                                : es::layernorm_t<__half>(x, ldx, out, ldo, gamma, beta, rows, c, eps, s);
}
