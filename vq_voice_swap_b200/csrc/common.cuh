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
// Fast APPROXIMATION of the erf-form GELU used by the CUDA-core prologues.  GELU(x) = x * Phi(x) with erfc from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): measured
// max |error| 4.2e-7 over [-12, 12] in fp32, tighter than ATen's own fp32 GELU (1.2e-6).
// Two MUFU ops (rcp, ex2) + 12 FP32 ops per element.
__device__ __forceinline__ float gelu_as(float x) {
  // GELU(x) = max(x, 0) - |x| * h(|x|),  h = 0.5*erfc(|x|/sqrt2) = (0.5*poly(t)) * t * exp(-x^2/2),
  // t = 1 / (1 + (0.3275911/sqrt2) |x|); the 0.5 and the 1/sqrt2 are folded into the constants.
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.2316418882f, ax, 1.0f));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float e = ex2_approx((x * x) * -0.72134752044f);
  const float h = (t * p) * e;
  return fmaf(-ax, h, fmaxf(x, 0.f));
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
