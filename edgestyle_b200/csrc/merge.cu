// EdgeStyle ControlNetBlock merge as three bandwidth passes with two per-sample global
// reductions.  The reference's stack/permute/contiguous interleave
// (/root/reference/model/edgestyle_multicontrolnet.py:160-164,479-514) is pure indexing here: group
// g = c*3+p of first_conv pairs nets (2p, 2p+1) of channel c, so the six residual slabs are read in place.
#include "common.h"
#include "ptx.cuh"

namespace es {

template <typename T>
__device__ __forceinline__ void mload8(const T* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = Cvt<T>::unpack2(u.x), b = Cvt<T>::unpack2(u.y), c = Cvt<T>::unpack2(u.z), d = Cvt<T>::unpack2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

struct MergeK {
  const void* res[6];
  float scale[6];
  int B, hw, C;
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  const void *g1, *be1, *g2, *be2;
  double* stats;
  float* z;
  const void* skip;
  long long lds;
  void* dst;
  long long ldd;
  float* gn_ws;
  int gn_groups, gn_cpg, gn_col0;
};

__device__ __forceinline__ void block_reduce2_to_global(float a, float b, double* dst) {
  __shared__ float red[2][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nw = (blockDim.x * blockDim.y + 31) >> 5;
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = a;
    red[1][tid >> 5] = b;
  }
  __syncthreads();
  if (tid < 32) {
    a = tid < nw ? red[0][tid] : 0.f;
    b = tid < nw ? red[1][tid] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (tid == 0) {
      atomicAdd(dst, static_cast<double>(a));
      atomicAdd(dst + 1, static_cast<double>(b));
    }
  }
}

// grid (chunks, B); block (C/8, ny). PHASE 1: stats of u. PHASE 2: z + stats of z. PHASE 3: output.
template <typename T, int PHASE>
__global__ void merge_kernel(const MergeK k) {
  pdl_launch_dependents();
  pdl_wait();  // PDL: inputs are produced by the preceding kernel
  const int v = threadIdx.x;
  const int ch = v * 8;
  const int b = blockIdx.y;
  const int C = k.C, hw = k.hw;
  const int per = (hw + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per;
  const int p1 = min(hw, p0 + per);
  const long long img_off = static_cast<long long>(b) * hw * C;

  float acc_s = 0.f, acc_q = 0.f;

  if (PHASE == 1 || PHASE == 2) {
    // per-channel first_conv weights with conditioning_scale folded in
    float wa[3][8], wb[3][8], bb[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        wa[p][j] = k.w1[((ch + j) * 3 + p) * 2 + 0] * k.scale[2 * p];
        wb[p][j] = k.w1[((ch + j) * 3 + p) * 2 + 1] * k.scale[2 * p + 1];
        bb[p][j] = k.b1[(ch + j) * 3 + p];
      }
    float mu1 = 0.f, r1 = 0.f;
    if (PHASE == 2) {
      const double n = 3.0 * C * hw;
      const double m = k.stats[b * 4 + 0] / n;
      const double var = k.stats[b * 4 + 1] / n - m * m;
      mu1 = static_cast<float>(m);
      r1 = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
    }
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
      const long long off = img_off + static_cast<long long>(p) * C + ch;
      float zacc[8];
      if (PHASE == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) zacc[j] = k.b2[ch + j];
      }
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        float ra[8], rb[8];
        mload8<T>(reinterpret_cast<const T*>(k.res[2 * pr]) + off, ra);
        mload8<T>(reinterpret_cast<const T*>(k.res[2 * pr + 1]) + off, rb);
        float u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = wa[pr][j] * ra[j] + wb[pr][j] * rb[j] + bb[pr][j];
        if (PHASE == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc_s += u[j];
            acc_q += u[j] * u[j];
          }
        } else {
          float g[8], be[8];
          const long long goff = (static_cast<long long>(p) * 3 + pr) * C + ch;
          mload8<T>(reinterpret_cast<const T*>(k.g1) + goff, g);
          mload8<T>(reinterpret_cast<const T*>(k.be1) + goff, be);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y = (u[j] - mu1) * r1 * g[j] + be[j];
            zacc[j] += k.w2[(ch + j) * 3 + pr] * silu_f(y);
          }
        }
      }
      if (PHASE == 2) {
        float* zp = k.z + off;
        reinterpret_cast<float4*>(zp)[0] = make_float4(zacc[0], zacc[1], zacc[2], zacc[3]);
        reinterpret_cast<float4*>(zp)[1] = make_float4(zacc[4], zacc[5], zacc[6], zacc[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc_s += zacc[j];
          acc_q += zacc[j] * zacc[j];
        }
      }
    }
    block_reduce2_to_global(acc_s, acc_q, k.stats + b * 4 + (PHASE == 1 ? 0 : 2));
  } else {
    const double n = 1.0 * C * hw;
    const double m = k.stats[b * 4 + 2] / n;
    const double var = k.stats[b * 4 + 3] / n - m * m;
    const float mu2 = static_cast<float>(m);
    const float r2 = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
    float w3[8], b3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w3[j] = k.w3[ch + j];
      b3[j] = k.b3[ch + j];
    }
    // optional GroupNorm statistics of the written rows (this thread's 8 channels touch at most two groups)
    __shared__ float gstat[66][2];
    const int g_lo = k.gn_ws ? (k.gn_col0 + ch) / k.gn_cpg : 0;
    const int g_split = k.gn_ws ? (g_lo + 1) * k.gn_cpg - (k.gn_col0 + ch) : 8;  // channels [0, g_split) -> g_lo
    const int g_base = k.gn_ws ? k.gn_col0 / k.gn_cpg : 0;
    float gs0 = 0.f, gq0 = 0.f, gs1 = 0.f, gq1 = 0.f;
    if (k.gn_ws) {
      for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 66 * 2; i += blockDim.x * blockDim.y) (&gstat[0][0])[i] = 0.f;
      __syncthreads();
    }
    for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
      const long long off = img_off + static_cast<long long>(p) * C + ch;
      const float4 z0 = reinterpret_cast<const float4*>(k.z + off)[0];
      const float4 z1 = reinterpret_cast<const float4*>(k.z + off)[1];
      const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      float g[8], be[8], sk[8];
      const long long goff = static_cast<long long>(p) * C + ch;
      mload8<T>(reinterpret_cast<const T*>(k.g2) + goff, g);
      mload8<T>(reinterpret_cast<const T*>(k.be2) + goff, be);
      const long long row = static_cast<long long>(b) * hw + p;
      if (k.skip) {
        mload8<T>(reinterpret_cast<const T*>(k.skip) + row * k.lds + ch, sk);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) sk[j] = 0.f;
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = sk[j] + w3[j] * silu_f((zz[j] - mu2) * r2 * g[j] + be[j]) + b3[j];
      if (k.gn_ws) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = Cvt<T>::to_f(Cvt<T>::from_f(o[j]));  // the rounded value the consumer will read
          if (j < g_split) { gs0 += v; gq0 += v * v; }
          else { gs1 += v; gq1 += v * v; }
        }
      }
      uint4 u;
      u.x = Cvt<T>::pack2(o[0], o[1]); u.y = Cvt<T>::pack2(o[2], o[3]);
      u.z = Cvt<T>::pack2(o[4], o[5]); u.w = Cvt<T>::pack2(o[6], o[7]);
      *reinterpret_cast<uint4*>(reinterpret_cast<T*>(k.dst) + row * k.ldd + ch) = u;
    }
    if (k.gn_ws) {
      atomicAdd(&gstat[g_lo - g_base][0], gs0);
      atomicAdd(&gstat[g_lo - g_base][1], gq0);
      if (g_split < 8) {
        atomicAdd(&gstat[g_lo - g_base + 1][0], gs1);
        atomicAdd(&gstat[g_lo - g_base + 1][1], gq1);
      }
      __syncthreads();
      for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 66 * 2; i += blockDim.x * blockDim.y) {
        const float v = (&gstat[0][0])[i];
        if (v != 0.f) atomicAdd(k.gn_ws + (static_cast<long long>(b) * k.gn_groups + g_base + (i >> 1)) * 2 + (i & 1), v);
      }
    }
  }
}

template <typename T>
static int merge_t(const EsMerge* m, int phase, cudaStream_t s) {
  MergeK k;
  for (int i = 0; i < 6; ++i) {
    k.res[i] = m->res[i];
    k.scale[i] = m->scale[i];
  }
  k.B = m->B; k.hw = m->hw; k.C = m->C;
  k.w1 = m->w1; k.b1 = m->b1; k.w2 = m->w2; k.b2 = m->b2; k.w3 = m->w3; k.b3 = m->b3;
  k.g1 = m->g1; k.be1 = m->be1; k.g2 = m->g2; k.be2 = m->be2;
  k.stats = m->stats; k.z = m->z; k.skip = m->skip; k.lds = m->lds; k.dst = m->dst; k.ldd = m->ldd;
  k.gn_ws = phase == 3 ? m->gn_ws : nullptr;
  k.gn_groups = m->gn_groups; k.gn_cpg = m->gn_cpg; k.gn_col0 = m->gn_col0;
  const int vpp = m->C / 8;
  int ny = 256 / vpp;
  if (ny < 1) ny = 1;
  if (ny > m->hw) ny = m->hw;
  dim3 block(vpp, ny, 1);
  int chunks = (4 * 148 + m->B - 1) / m->B;
  const int max_chunks = (m->hw + ny * 2 - 1) / (ny * 2);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, m->B, 1);
  if (phase == 1) ES_CUDA(launch_kernel(merge_kernel<T, 1>, dim3(grid), dim3(block), 0, s, k));
  else if (phase == 2) ES_CUDA(launch_kernel(merge_kernel<T, 2>, dim3(grid), dim3(block), 0, s, k));
  else ES_CUDA(launch_kernel(merge_kernel<T, 3>, dim3(grid), dim3(block), 0, s, k));
  ES_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace es

extern "C" int es_merge_phase(const EsMerge* m, int phase, void* stream) {
  ES_CHECK(m && phase >= 1 && phase <= 3, "es_merge_phase: bad arguments");
  ES_CHECK(m->C % 8 == 0 && m->C / 8 <= 1024 && m->stats && m->z, "es_merge_phase: bad shape/workspace");
  if (phase == 3) ES_CHECK(m->dst && m->ldd % 8 == 0 && (!m->skip || m->lds % 8 == 0), "es_merge_phase: bad dst");
  if (phase == 3 && m->gn_ws)
    ES_CHECK(m->gn_cpg >= 8 && m->gn_col0 >= 0 && m->gn_groups > 0 && m->C / m->gn_cpg + 2 <= 66 &&
                 (m->gn_col0 + m->C + m->gn_cpg - 1) / m->gn_cpg <= m->gn_groups,
             "es_merge_phase: bad GroupNorm slice");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return m->dtype == ES_DTYPE_BF16 ? es::merge_t<__nv_bfloat16>(m, phase, s) : es::merge_t<__half>(m, phase, s);
}
