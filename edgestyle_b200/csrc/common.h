// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/edgestyle_b200.h"

namespace es {

void set_error(const char* fmt, ...);

#define ES_CHECK(cond, ...)        \
  do {                             \
    if (!(cond)) {                 \
      ::es::set_error(__VA_ARGS__); \
      return -1;                   \
    }                              \
  } while (0)

#define ES_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::es::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1. 16-bit elements, SWIZZLE_128B.
int encode_tmap_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool swizzle128 = true);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL): when enabled, every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be scheduled (and run their prologue:
// barrier init, TMEM alloc, tensor-map prefetch) while the previous kernel in the stream drains; each kernel
// executes griddepcontrol.wait before its first global-memory access.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int n = 0;
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    n = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace es
