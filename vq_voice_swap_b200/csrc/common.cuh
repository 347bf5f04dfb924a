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
// erf-form GELU through libm's erff, reference models/unet.py:341-342 (nn.GELU default).
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Fast APPROXIMATION of the erf-form GELU used by the CUDA-core prologues (the tcgen05 prologue has the packed fp32x2 twin,
// umma_conv.cu gelu4p).  GELU(x) = max(x, 0) - |x| * Phi(-|x|) with Phi(-a) = 2^P5(a), P5 the weighted-minimax fit of
// log2 Phi(-a) (tools/fit_gelu_poly.py): |error| <= 4.4e-7 exact, 6.4e-7 as evaluated in fp32 (ATen's own fp32 GELU: 1.2e-6),
// monotone for all a, so large |x| flush to 0 through ex2(-inf).  One MUFU op + 8 FP32 ops per element.
__device__ __forceinline__ float gelu_as(float x) {
  const float n = -fabsf(x);  // Horner in n = -|x|: the odd coefficients of P5 change sign
  float p = fmaf(4.733092792e-04f, n, 7.084557321e-03f);
  p = fmaf(p, n, 5.182738230e-02f);
  p = fmaf(p, n, -4.599924386e-01f);
  p = fmaf(p, n, 1.150787830e+00f);
  p = fmaf(p, n, -1.000037670e+00f);
  return fmaf(n, ex2_approx(p), fmaxf(x, 0.f));
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
