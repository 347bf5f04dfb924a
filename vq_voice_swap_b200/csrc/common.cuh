// Shared host/device helpers for libvqvs (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqvs.h"

namespace vqvs {

// ---- host-side error channel (thread-local, see vqvs_last_error) -------------
void set_error(const char* fmt, ...);

#define VQVS_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::vqvs::set_error(__VA_ARGS__);      \
      return VQVS_EINVAL;                  \
    }                                      \
  } while (0)

#define VQVS_CHECK_LAUNCH(what)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::vqvs::set_error("%s: CUDA error: %s", what, cudaGetErrorString(e__));     \
      return VQVS_ECUDA;                                                          \
    }                                                                             \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------
// Exact-erf GELU, reference models/unet.py:341-342 (nn.GELU default).
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// PyTorch's nearest-neighbour source index (F.interpolate(x, size), reference
// models/unet.py:139): computed in fp32 as floor(dst * (in/out)), clamped.
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

}  // namespace vqvs
