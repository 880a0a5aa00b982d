// Shared host-side helpers for the C-ABI library: error reporting, launch checks, device info.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/imgenh_b200.h"

namespace ie {

void set_error(const char* fmt, ...);
int sm_count();

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
// rank-2 bf16 tensor [rows][cols] with `pitch_elems` elements between rows; box = box_cols x box_rows;
// 128-byte swizzle (box_cols must be 64).  Returns 0 or <0 with the error set.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                      uint32_t box_cols, uint32_t box_rows);

// rank-4 bf16 NHWC tensor in TMA im2col mode ('same' convolution with `pad` zero pixels); see api.cu
int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, int n, int h, int w, uint64_t pitch_elems, int pad,
                          uint32_t pixels);

// Launch with the programmatic-stream-serialization attribute (see pdl_wait in ie_ptx.cuh).  The kernel MUST call
// pdl_wait() before touching anything an earlier kernel of the stream produced.  IE_PDL=0 in the environment (or
// ie_conv_set_mode flag bit 10) turns the attribute off: plain stream-ordered launches.
bool pdl_enabled();
void pdl_set(bool on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ie

#define IE_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ie::set_error(__VA_ARGS__);    \
      return IE_ERR_INVALID;         \
    }                                \
  } while (0)

#define IE_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ie::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return IE_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define IE_LAUNCH_CHECK() IE_CUDA(cudaGetLastError())

static inline int ie_ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }
