// Micro-benchmark + numerics probe for tcgen05.mma shared-memory operand layouts (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mma_bench.cu -o tools/mma_bench
// Answers two questions for the conv kernel's operand staging:
//   1. cycles per MMA (M=128, N, K=16 bf16) for the K-major layouts NONE / 32B / 64B / 128B swizzle,
//      with row-shifted start addresses (the conv taps);
//   2. whether a swizzled tile read from a row-shifted start address still yields the right numbers,
//      and whether the descriptor's base_offset field has to carry the shift.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t holder, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// layout: 0 none, 6 = 32B, 4 = 64B, 2 = 128B swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout, int base_off) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ int row_bytes(int layout) { return layout == 2 ? 128 : layout == 4 ? 64 : layout == 6 ? 32 : 16; }

// ---------------------------------------------------------------------------------------------
// timing: thread 0 issues `iters` groups of (4 K blocks x 3 taps x 3 products) MMAs
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bench_kernel(int layout, int n, int shift_rows, int iters, int busy, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = holder;
  const int rb = row_bytes(layout);
  const int rows = 128 + 2 * 32;
  // A: hi tile then lo tile; per layout the K extent of one row is rb bytes (rb/32 K blocks per tile, min 1)
  const uint32_t a_base = smem_u32(smem);
  const uint32_t tile_a = rows * (layout ? rb : 64);   // none: [2 chunks][rows][16B] per K block
  const uint32_t b_base = a_base + 96 * 1024;
  const uint32_t tile_b = n * (layout ? rb : 64);
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kb = 0; kb < 4; ++kb) {
        for (int tap = 0; tap < 3; ++tap) {
          uint32_t a_hi, a_lo, b_hi, b_lo;
          uint64_t da_hi, da_lo, db_hi, db_lo;
          if (layout == 0) {
            a_hi = a_base + kb * (rows * 64) + tap * shift_rows * 16; a_lo = a_hi + rows * 32;
            b_hi = b_base + ((kb * 3 + tap) * (n * 64)) % (32 * 1024); b_lo = b_hi + n * 32;
            da_hi = make_desc(a_hi, rows * 16, 128, 0, 0); da_lo = make_desc(a_lo, rows * 16, 128, 0, 0);
            db_hi = make_desc(b_hi, n * 16, 128, 0, 0); db_lo = make_desc(b_lo, n * 16, 128, 0, 0);
          } else {
            const int kpr = rb / 32;  // K blocks per row
            a_hi = a_base + (kb / kpr) * 2 * tile_a + (kb % kpr) * 32 + tap * shift_rows * rb; a_lo = a_hi + tile_a;
            b_hi = b_base + (((kb / kpr) * 3 + tap) * 2 * tile_b) % (32 * 1024) + (kb % kpr) * 32; b_lo = b_hi + tile_b;
            da_hi = make_desc(a_hi, 16, 8 * rb, layout, 0); da_lo = make_desc(a_lo, 16, 8 * rb, layout, 0);
            db_hi = make_desc(b_hi, 16, 8 * rb, layout, 0); db_lo = make_desc(b_lo, 16, 8 * rb, layout, 0);
          }
          mma_bf16(tmem, da_hi, db_hi, idesc, 1);
          mma_bf16(tmem, da_lo, db_hi, idesc, 1);
          mma_bf16(tmem, da_hi, db_lo, idesc, 1);
        }
      }
    }
    mma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (busy && warp >= 1) {
    // competing LDS/STS traffic from 3 warps while the MMAs run
    const uint32_t p = smem_u32(smem + 160 * 1024);
    uint32_t x = 0, y = 1, z = 2, w = 3;
    for (int i = 0; i < busy; ++i) {
      uint32_t a0, a1, a2, a3;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(p + (((threadIdx.x + i * 96) & 1023) << 4)));
      x ^= a0; y ^= a1; z ^= a2; w ^= a3;
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(p + (((threadIdx.x * 7 + i) & 1023) << 4)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
    }
    if (x == 0x12345) out[200] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_dealloc(tmem, 512);
  }
}


__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t make_idesc_m(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// Clean MMA-rate probe: the whole warp runs the loop with warp-uniform descriptors (no R2UR waterfall), one elected
// lane issues.  `group` MMAs are issued back to back per commit; ndiff = number of distinct accumulators cycled through.
__global__ void __launch_bounds__(128) bench2_kernel(int layout, int m, int n, int iters, int group, int ndiff, long long* out, int ashift = 0, int bshift = 0, int contend = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = holder;
  if (warp == 0) {
    const int rb = layout == 2 ? 128 : 16;
    const uint32_t a_base = smem_u32(smem) >> 4, b_base = (smem_u32(smem) + 96 * 1024) >> 4;
    const uint32_t idesc = make_idesc_m(m, n);
    const uint64_t a_const = layout ? make_desc(0, 16, 8 * rb, layout, 0) : make_desc(0, (128 + 8) * 16, 128, 0, 0);
    const uint64_t b_const = layout ? make_desc(0, 16, 8 * rb, layout, 0) : make_desc(0, n * 16, 128, 0, 0);
    const long long t0 = clock64();
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        const uint64_t da0 = a_const + (a_base + ashift), da1 = a_const + (a_base + 544 + 2 * ashift), db0 = b_const + (b_base + bshift), db1 = b_const + (b_base + 1024 + bshift);
        const uint32_t d1 = tmem + (ndiff > 1 ? (uint32_t)n : 0u);
#pragma unroll 1
        for (int k = 0; k < group; k += 4) {  // pure back-to-back issue: descriptors are loop-invariant
          mma_bf16(tmem, da0, db0, idesc, 1);
          mma_bf16(d1, da1, db0, idesc, 1);
          mma_bf16(tmem, da0, db1, idesc, 1);
          mma_bf16(d1, da1, db1, idesc, 1);
        }
        mma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), par);
      par ^= 1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    *reinterpret_cast<volatile uint32_t*>(&holder) = 0xffffffffu;  // stop the contending warps
  } else if (contend == 1) {
    uint32_t acc = 0;
    while (*reinterpret_cast<volatile uint32_t*>(&holder) != 0xffffffffu) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + 256 + (acc & 127)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r[0] + 16;
    }
    if (acc == 0x1234567) out[300] = 1;
  } else if (contend == 2) {
    const uint32_t p = smem_u32(smem + 160 * 1024);
    uint32_t x = 0, y = 1, z = 2, w = 3, i = 0;
    while (*reinterpret_cast<volatile uint32_t*>(&holder) != 0xffffffffu) {
      uint32_t a0, a1, a2, a3;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(p + (((threadIdx.x + i * 96) & 1023) << 4)));
      x ^= a0; y ^= a1; z ^= a2; w ^= a3;
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(p + (((threadIdx.x * 7 + i) & 1023) << 4)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
      ++i;
    }
    if (x == 0x12345) out[200] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// numerics: D[128, n] = A[shift .. shift+128, 0..k) * B[n, k]^T with bf16-exact inputs
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swz(uint32_t off, int layout) {
  if (layout == 2) return off ^ (((off >> 7) & 7) << 4);
  if (layout == 4) return off ^ (((off >> 7) & 3) << 4);
  if (layout == 6) return off ^ (((off >> 7) & 1) << 4);
  return off;
}

__global__ void __launch_bounds__(128) numerics_kernel(int layout, int n, int k, int shift, int base_mode, const float* a,
                                                       const float* b, float* dout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = row_bytes(layout), kpr = rb / 2;  // K elements per row
  const int rows = 128 + shift;
  const int ktiles = (k + kpr - 1) / kpr;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 96 * 1024;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 256);
  const uint32_t a_tile = ((rows * rb + 1023) / 1024) * 1024, b_tile = ((n * rb + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < rows * k; i += 128) {
    const int r = i / k, c = i % k;
    const uint32_t off = r * rb + (c % kpr) * 2;
    *reinterpret_cast<__nv_bfloat16*>(a_s + (c / kpr) * a_tile + swz(off, layout)) = __float2bfloat16_rn(a[(size_t)r * k + c]);
  }
  for (int i = threadIdx.x; i < n * k; i += 128) {
    const int r = i / k, c = i % k;
    const uint32_t off = r * rb + (c % kpr) * 2;
    *reinterpret_cast<__nv_bfloat16*>(b_s + (c / kpr) * b_tile + swz(off, layout)) = __float2bfloat16_rn(b[(size_t)r * k + c]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = holder;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    uint32_t acc = 0;
    for (int kb = 0; kb < k / 16; ++kb) {
      const int per = kpr / 16;  // K blocks per row
      const uint32_t a_addr = smem_u32(a_s + (kb / per) * a_tile) + (kb % per) * 32 + shift * rb;
      const uint32_t b_addr = smem_u32(b_s + (kb / per) * b_tile) + (kb % per) * 32;
      const int bo = base_mode == 1 ? (int)((a_addr >> 7) & 7) : base_mode == 2 ? (shift & 7) : 0;
      mma_bf16(tmem, make_desc(a_addr, 16, 8 * rb, layout, bo), make_desc(b_addr, 16, 8 * rb, layout, 0), idesc, acc);
      acc = 1;
    }
    mma_commit(smem_u32(&bar));
  }
  (void)ktiles;
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) dout[(size_t)row * n + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_dealloc(tmem, 256);
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(numerics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* d_out;
  CK(cudaMalloc(&d_out, 512 * sizeof(long long)));
  const int layouts[4] = {0, 6, 4, 2};
  const char* names[4] = {"none", "sw32", "sw64", "sw128"};
  CK(cudaFuncSetAttribute(bench2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("== clean MMA rate (warp-uniform descriptors, elected issue): cycles per MMA incl. one commit+wait per group ==\n");
  for (int layout : {0, 2})
    for (int m : {128, 64})
      for (int n : {64, 128, 256})
        for (int group : {12, 96})
          for (int ndiff : {1, 2}) {
            if (ndiff * n > 512) continue;
            const int iters = 200;
            bench2_kernel<<<148, 128, 200 * 1024>>>(layout, m, n, iters, group, ndiff, d_out);
            CK(cudaDeviceSynchronize());
            long long h[148];
            CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += (double)h[i];
            avg /= 148;
            printf("layout=%-5s M=%3d N=%3d group=%2d accumulators=%d : %.1f cycles/MMA\n", layout ? "sw128" : "none", m, n, group, ndiff,
                   avg / (iters * (double)group));
          }
  printf("== effect of 16-byte-granular operand start shifts (conv taps) at the true pipe rate, layout none, M=128 ==\n");
  for (int n : {64, 128, 256})
    for (int ashift : {0, 1, 2, 4, 8})
      for (int bshift : {0, 1}) {
        const int iters = 200, group = 96;
        bench2_kernel<<<148, 128, 200 * 1024>>>(0, 128, n, iters, group, 1, d_out, ashift, bshift);
        CK(cudaDeviceSynchronize());
        long long h[148];
        CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148;
        printf("N=%3d A start +%d x16B, B start +%d x16B : %.1f cycles/MMA\n", n, ashift, bshift, avg / (iters * (double)group));
      }
  printf("== contention: 3 other warps running tcgen05.ld (1) or LDS/STS.128 (2) while the MMAs run; layout none, M=128 ==\n");
  for (int n : {64, 128, 256})
    for (int contend : {0, 1, 2}) {
      const int iters = 200, group = 96;
      bench2_kernel<<<148, 128, 200 * 1024>>>(0, 128, n, iters, group, 1, d_out, 0, 0, contend);
      CK(cudaDeviceSynchronize());
      long long h[148];
      CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      avg /= 148;
      printf("N=%3d contend=%d : %.1f cycles/MMA\n", n, contend, avg / (iters * (double)group));
    }
  if (getenv("MMA_BENCH_OLD")) {
  printf("== cycles per MMA (M=128, K=16, bf16), 36 MMAs per group, thread-0 issue -> commit -> wait ==\n");
  for (int grid : {1, 148}) {
    for (int busy : {0, 4000}) {
      for (int li = 0; li < 4; ++li) {
        for (int n : {64, 128, 256}) {
          for (int shift : {0, 1, 2}) {
            const int iters = 50;
            bench_kernel<<<grid, 128, 200 * 1024>>>(layouts[li], n, shift, iters, busy, d_out);
            CK(cudaDeviceSynchronize());
            long long h[148];
            CK(cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            double avg = 0;
            for (int i = 0; i < grid; ++i) avg += (double)h[i];
            avg /= grid;
            printf("grid=%3d busy=%d layout=%-5s n=%3d tap_shift=%d : %.1f cycles/MMA\n", grid, busy ? 1 : 0, names[li], n, shift,
                   avg / (iters * 36.0));
          }
        }
      }
    }
  }
  }
  if (!getenv("MMA_BENCH_OLD")) return 0;
  printf("== numerics: row-shifted start address under swizzle ==\n");
  const int n = 64, k = 64;
  for (int li = 0; li < 4; ++li) {
    if (layouts[li] == 0) continue;
    for (int shift : {0, 1, 2, 3, 4, 8, 32, 33}) {
      for (int base_mode = 0; base_mode < 3; ++base_mode) {
        const int rows = 128 + shift;
        std::vector<float> a((size_t)rows * k), b((size_t)n * k), ref((size_t)128 * n), got((size_t)128 * n);
        for (size_t i = 0; i < a.size(); ++i) a[i] = (float)((int)((i * 2654435761u) >> 27) - 16);
        for (size_t i = 0; i < b.size(); ++i) b[i] = (float)((int)((i * 40503u + 7) % 17) - 8);
        for (int r = 0; r < 128; ++r)
          for (int c = 0; c < n; ++c) {
            float s = 0;
            for (int x = 0; x < k; ++x) s += a[(size_t)(r + shift) * k + x] * b[(size_t)c * k + x];
            ref[(size_t)r * n + c] = s;
          }
        float *da, *db, *dd;
        CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dd, got.size() * 4));
        CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
        numerics_kernel<<<1, 128, 200 * 1024>>>(layouts[li], n, k, shift, base_mode, da, db, dd);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (size_t i = 0; i < got.size(); ++i) maxerr = fmax(maxerr, fabs((double)got[i] - ref[i]));
        printf("layout=%-5s shift=%2d base_mode=%d : max|err| = %.1f %s\n", names[li], shift, base_mode, maxerr, maxerr == 0 ? "OK" : "WRONG");
        cudaFree(da); cudaFree(db); cudaFree(dd);
      }
    }
  }
  return 0;
}
