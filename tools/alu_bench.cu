// Throughput probes for the transform inner loop (sm_100a): FFMA vs FFMA2, MUFU, and the two GELU formulations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/alu_bench.cu -o tools/alu_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t bcast2(float c) { return pack2(c, c); }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_as(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.2316418882f, ax, 1.0f));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f); p = fmaf(p, t, -0.142248368f); p = fmaf(p, t, 0.127414796f);
  const float e = ex2_approx((x * x) * -0.72134752044f);
  const float h = (t * p) * e;
  return fmaf(-ax, h, fmaxf(x, 0.f));
}
__device__ __forceinline__ void gelu4p(uint64_t* y) {
  uint64_t nay[4], t[4], q[4], e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) nay[i] = y[i] | 0x8000000080000000ull;
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = fma2(nay[i], bcast2(-0.2316418882f), bcast2(1.0f));
#pragma unroll
  for (int i = 0; i < 4; ++i) { float a, b; unpack2(t[i], a, b); t[i] = pack2(rcp_approx(a), rcp_approx(b)); }
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(y[i], y[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = mul2(e[i], bcast2(-0.72134752044f));
#pragma unroll
  for (int i = 0; i < 4; ++i) { float a, b; unpack2(e[i], a, b); e[i] = pack2(ex2_approx(a), ex2_approx(b)); }
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(t[i], bcast2(0.5307027145f), bcast2(-0.7265760135f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], bcast2(0.7107068705f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], bcast2(-0.142248368f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = fma2(q[i], t[i], bcast2(0.127414796f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = mul2(q[i], t[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = mul2(q[i], e[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float a, b; unpack2(y[i], a, b); y[i] = fma2(nay[i], q[i], pack2(fmaxf(a, 0.f), fmaxf(b, 0.f))); }
}

// mode 0: 8 independent FFMA chains; 1: 8 independent FFMA2 chains; 2: MUFU.RCP x8; 3: MUFU.EX2 x8;
// 4: scalar gelu on 8 values; 5: packed gelu4p on 8 values
__global__ void bench(int mode, int iters, float seed, float* out, long long* cyc) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.01f * (threadIdx.x + i);
  uint64_t p[4];
  for (int i = 0; i < 4; ++i) p[i] = pack2(v[2 * i], v[2 * i + 1]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(seed), "f"(0.5f));
    } else if (mode == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], bcast2(seed), bcast2(0.5f));
    } else if (mode == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = rcp_approx(v[i]);
    } else if (mode == 3) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ex2_approx(v[i]);
    } else if (mode == 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_as(v[i] * seed + 0.3f);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], bcast2(seed), bcast2(0.3f));
      gelu4p(p);
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  for (int i = 0; i < 4; ++i) { float a, b; unpack2(p[i], a, b); s += a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const char* names[6] = {"FFMA x32 (8 chains)", "FFMA2 x16 (4 chains)", "MUFU.RCP x32", "MUFU.EX2 x32", "gelu scalar x8", "gelu packed x8"};
  const int per_iter[6] = {32, 16, 32, 32, 8, 8};
  for (int mode = 0; mode < 6; ++mode)
    for (int threads : {32, 128, 256, 512, 1024}) {
      const int iters = 2000;
      bench<<<148, threads>>>(mode, iters, 0.999f, out, cyc);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      const double warps = threads / 32.0;
      printf("%-22s warps/SM=%4.0f: %.2f cycles per iteration per warp, %.3f cycles per warp-instr-or-elem at SM level (= per SMSP x4: %.2f)\n",
             names[mode], warps, avg / iters, avg / iters / per_iter[mode] / warps, 4 * avg / iters / per_iter[mode] / warps);
    }
  return 0;
}
