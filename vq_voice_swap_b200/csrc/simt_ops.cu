// CUDA-core (fp32 SIMT) kernels of libvqvs: everything on the hot path that is not a
// wide convolution, plus a fully general fused conv used for odd shapes and as the
// on-device cross-check of the tcgen05 path (umma_conv.cu).
#include "common.cuh"

namespace vqvs {

// =============================================================================
// GroupNorm finalize: channel (sum, sumsq) -> per-(n,c) scale/shift (+FiLM)
// =============================================================================
__global__ void gn_finalize_kernel(VqvsGnFinalize d) {
  const int C = d.c_a + d.c_b;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.batch * C) return;
  const int n = idx / C, c = idx - n * C;
  const int cg = C / d.groups;
  const int g = c / cg;
  double s = 0.0, ss = 0.0;
  for (int cc = g * cg; cc < (g + 1) * cg; ++cc) {
    const double* st = cc < d.c_a ? d.stats_a + ((size_t)n * d.c_a + cc) * 2
                                  : d.stats_b + ((size_t)n * d.c_b + (cc - d.c_a)) * 2;
    s += st[0];
    ss += st[1];
  }
  const double cnt = (double)cg * (double)d.count;
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double rstd = rsqrt(var + 1e-5);
  double sc = rstd * (double)d.gamma[c];
  double sh = (double)d.beta[c] - mean * sc;
  if (d.film) {
    const double a = (double)d.film[(size_t)n * d.film_stride + c];
    const double b = (double)d.film[(size_t)n * d.film_stride + C + c];
    sc = sc * (1.0 + a);
    sh = sh * (1.0 + a) + b;
  }
  d.scale[idx] = (float)sc;
  d.shift[idx] = (float)sh;
}

// =============================================================================
// standalone per-channel statistics
// =============================================================================
__global__ void channel_stats_kernel(const float* __restrict__ x, int t, double* stats) {
  const size_t row = blockIdx.x;  // n*C + c
  const float* p = x + row * (size_t)t;
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < t; i += blockDim.x) {
    const double v = (double)p[i];
    s += v;
    ss += v * v;
  }
  __shared__ double red[2][32];
  s = warp_sum(s);
  ss = warp_sum(ss);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = s; red[1][w] = ss; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    s = l < nw ? red[0][l] : 0.0;
    ss = l < nw ? red[1][l] : 0.0;
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (l == 0) {
      atomicAdd(stats + row * 2, s);
      atomicAdd(stats + row * 2 + 1, ss);
    }
  }
}

// =============================================================================
// in_conv: Conv1d(1 -> C, k3, pad 1) (+ nearest-upsampled conditioning) + stats
// =============================================================================
constexpr int CIN_THREADS = 256;
constexpr int CIN_PER_THREAD = 4;

__global__ void __launch_bounds__(CIN_THREADS) conv_in_kernel(VqvsConvIn d) {
  extern __shared__ float cin_smem[];  // [8 warps][c_out][2]
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * (CIN_THREADS * CIN_PER_THREAD);
  const float* x = d.x + (size_t)n * d.t;
  float xm[CIN_PER_THREAD], xc[CIN_PER_THREAD], xp[CIN_PER_THREAD];
  int csrc[CIN_PER_THREAD];
  bool valid[CIN_PER_THREAD];
  const float cscale = d.cond ? (float)d.t_cond / (float)d.t : 0.f;
#pragma unroll
  for (int j = 0; j < CIN_PER_THREAD; ++j) {
    const int t = t0 + threadIdx.x + j * CIN_THREADS;
    valid[j] = t < d.t;
    xm[j] = (valid[j] && t - 1 >= 0) ? x[t - 1] : 0.f;
    xc[j] = valid[j] ? x[t] : 0.f;
    xp[j] = (valid[j] && t + 1 < d.t) ? x[t + 1] : 0.f;
    csrc[j] = d.cond ? nearest_src(t, cscale, d.t_cond) : 0;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = 0; o < d.c_out; ++o) {
    const float w0 = d.w[o * 3], w1 = d.w[o * 3 + 1], w2 = d.w[o * 3 + 2], b = d.bias[o];
    float* out = d.out + ((size_t)n * d.c_out + o) * d.t;
    const float* cond = d.cond ? d.cond + ((size_t)n * d.c_out + o) * d.t_cond : nullptr;
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < CIN_PER_THREAD; ++j) {
      if (valid[j]) {
        float v = b + w0 * xm[j] + w1 * xc[j] + w2 * xp[j];
        if (cond) v += cond[csrc[j]];
        out[t0 + threadIdx.x + j * CIN_THREADS] = v;
        s += v;
        ss += v * v;
      }
    }
    if (d.stats_out) {
      s = warp_sum(s);
      ss = warp_sum(ss);
      if (lane == 0) {
        cin_smem[(warp * d.c_out + o) * 2] = s;
        cin_smem[(warp * d.c_out + o) * 2 + 1] = ss;
      }
    }
  }
  if (d.stats_out) {
    __syncthreads();
    for (int i = threadIdx.x; i < d.c_out * 2; i += CIN_THREADS) {
      double acc = 0.0;
      for (int w = 0; w < CIN_THREADS / 32; ++w) acc += (double)cin_smem[w * d.c_out * 2 + i];
      atomicAdd(d.stats_out + (size_t)n * d.c_out * 2 + i, acc);
    }
  }
}

// =============================================================================
// out conv: GN-affine -> GELU -> Conv1d(C -> 1, k3) with the DDPM update fused
// =============================================================================
constexpr int COUT_THREADS = 256;
constexpr int COUT_VEC = 4;  // positions per thread of the vectorised kernel

// DDPM store of one output position (shared by both conv_out kernels); returns the x0 term of mode X0_SUM.
__device__ __forceinline__ double conv_out_store(const VqvsConvOut& d, const float* cf, size_t o, float eps) {
  if (d.mode == VQVS_OUT_PREV) {
    const float xt = d.x_t[o];
    // alpha^-1/2 * (x_t - (beta*(1-abar)^-1/2) * eps) + sigma * noise, unfused like the reference
    float v = __fmul_rn(cf[0], __fsub_rn(xt, __fmul_rn(cf[1], eps)));
    if (d.noise) v = __fadd_rn(v, __fmul_rn(cf[2], d.noise[o]));
    d.out[o] = v;
    return 0.0;
  }
  d.out[o] = eps;
  if (d.mode == VQVS_OUT_X0_SUM) return (double)__fmul_rn(__fsub_rn(d.x_t[o], __fmul_rn(cf[3], eps)), cf[4]);
  return 0.0;
}

// Vectorised variant (t % 4 == 0): every thread owns 4 consecutive positions, reads them with one 128-bit load per
// channel, evaluates GN-affine + GELU once per element and gets the two halo values from its neighbour lanes.
__global__ void __launch_bounds__(COUT_THREADS) conv_out_vec_kernel(VqvsConvOut d) {
  __shared__ double red[COUT_THREADS / 32];
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * COUT_THREADS + threadIdx.x) * COUT_VEC;
  const bool valid = t0 < d.t;  // t % 4 == 0: a thread's 4 positions are all inside or all outside
  float acc[COUT_VEC] = {0.f, 0.f, 0.f, 0.f};
  const float* xn = d.x + (size_t)n * d.c_in * d.t;
  for (int c = 0; c < d.c_in; ++c) {
    const float sc = __ldg(d.scale + n * d.c_in + c), sh = __ldg(d.shift + n * d.c_in + c);
    const float* x = xn + (size_t)c * d.t;
    float g[COUT_VEC] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + t0));
      g[0] = gelu_as(fmaf(v.x, sc, sh));
      g[1] = gelu_as(fmaf(v.y, sc, sh));
      g[2] = gelu_as(fmaf(v.z, sc, sh));
      g[3] = gelu_as(fmaf(v.w, sc, sh));
    }
    float left = __shfl_up_sync(0xffffffffu, g[3], 1), right = __shfl_down_sync(0xffffffffu, g[0], 1);
    if (lane == 0) left = (valid && t0 > 0) ? gelu_as(fmaf(__ldg(x + t0 - 1), sc, sh)) : 0.f;
    if (lane == 31) right = (valid && t0 + COUT_VEC < d.t) ? gelu_as(fmaf(__ldg(x + t0 + COUT_VEC), sc, sh)) : 0.f;
    const float w0 = __ldg(d.w + c * 3), w1 = __ldg(d.w + c * 3 + 1), w2 = __ldg(d.w + c * 3 + 2);
    acc[0] += w0 * left + w1 * g[0] + w2 * g[1];
    acc[1] += w0 * g[0] + w1 * g[1] + w2 * g[2];
    acc[2] += w0 * g[1] + w1 * g[2] + w2 * g[3];
    acc[3] += w0 * g[2] + w1 * g[3] + w2 * right;
  }
  const float* cf = d.coef ? d.coef + n * 8 : nullptr;
  double x0 = 0.0;
  if (valid) {
    const float bias = d.bias[0];
#pragma unroll
    for (int i = 0; i < COUT_VEC; ++i) x0 += conv_out_store(d, cf, (size_t)n * d.t + t0 + i, acc[i] + bias);
  }
  if (d.mode == VQVS_OUT_X0_SUM) {
    x0 = warp_sum(x0);
    if (lane == 0) red[threadIdx.x >> 5] = x0;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < COUT_THREADS / 32; ++w) s += red[w];
      atomicAdd(d.x0_sum + n, s);
    }
  }
}

__global__ void __launch_bounds__(COUT_THREADS) conv_out_kernel(VqvsConvOut d) {
  __shared__ float row[2][COUT_THREADS + 2];
  __shared__ double red[COUT_THREADS / 32];
  const int n = blockIdx.y;
  const int t = blockIdx.x * COUT_THREADS + threadIdx.x;
  const bool valid = t < d.t;
  float acc = 0.f;
  for (int c = 0; c < d.c_in; ++c) {
    const float sc = d.scale[n * d.c_in + c], sh = d.shift[n * d.c_in + c];
    const float* x = d.x + ((size_t)n * d.c_in + c) * d.t;
    float* r = row[c & 1];
    r[threadIdx.x + 1] = valid ? gelu_erf(x[t] * sc + sh) : 0.f;
    if (threadIdx.x == 0) {
      const int tl = blockIdx.x * COUT_THREADS - 1;
      r[0] = tl >= 0 ? gelu_erf(x[tl] * sc + sh) : 0.f;
    } else if (threadIdx.x == COUT_THREADS - 1) {
      const int tr = blockIdx.x * COUT_THREADS + COUT_THREADS;
      r[COUT_THREADS + 1] = tr < d.t ? gelu_erf(x[tr] * sc + sh) : 0.f;
    }
    __syncthreads();
    const float w0 = d.w[c * 3], w1 = d.w[c * 3 + 1], w2 = d.w[c * 3 + 2];
    acc += w0 * r[threadIdx.x] + w1 * r[threadIdx.x + 1] + w2 * r[threadIdx.x + 2];
  }
  const float* cf = d.coef ? d.coef + n * 8 : nullptr;
  double x0 = 0.0;
  if (valid) x0 = conv_out_store(d, cf, (size_t)n * d.t + t, acc + d.bias[0]);
  if (d.mode == VQVS_OUT_X0_SUM) {
    x0 = warp_sum(x0);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x0;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < COUT_THREADS / 32; ++w) s += red[w];
      atomicAdd(d.x0_sum + n, s);
    }
  }
}

// =============================================================================
// DDPM elementwise finisher and x0 partial sums
// =============================================================================
__global__ void ddpm_finish_kernel(VqvsDdpmFinish d) {
  const int n = blockIdx.y;
  const float* cf = d.coef + n * 8;
  const float mean = d.use_x0_mean ? (float)(d.x0_sum[n] / (double)d.t) : 0.f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < d.t; t += gridDim.x * blockDim.x) {
    const size_t o = (size_t)n * d.t + t;
    const float xt = d.x_t[o];
    float eps = d.eps[o];
    if (d.use_x0_mean) {
      float x0 = __fmul_rn(__fsub_rn(xt, __fmul_rn(cf[3], eps)), cf[4]);
      x0 = fminf(fmaxf(__fsub_rn(x0, mean), -1.f), 1.f);
      eps = __fmul_rn(__fsub_rn(xt, __fmul_rn(x0, cf[5])), cf[6]);
    }
    float v = __fmul_rn(cf[0], __fsub_rn(xt, __fmul_rn(cf[1], eps)));
    if (d.noise) v = __fadd_rn(v, __fmul_rn(cf[2], d.noise[o]));
    d.out[o] = v;
  }
}

__global__ void ddpm_x0_sum_kernel(const float* __restrict__ x_t, const float* __restrict__ eps,
                                   const float* __restrict__ coef, int t_len, double* x0_sum) {
  const int n = blockIdx.y;
  const float c3 = coef[n * 8 + 3], c4 = coef[n * 8 + 4];
  double s = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < t_len; t += gridDim.x * blockDim.x) {
    const size_t o = (size_t)n * t_len + t;
    s += (double)__fmul_rn(__fsub_rn(x_t[o], __fmul_rn(c3, eps[o])), c4);
  }
  __shared__ double red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w];
    atomicAdd(x0_sum + n, a);
  }
}

// =============================================================================
// timestep embedding MLP and the stacked FiLM Linear
// =============================================================================
__global__ void __launch_bounds__(256) time_embed_kernel(VqvsTimeEmbed d) {
  extern __shared__ float te_smem[];  // e[dim], h[dim]
  float* e = te_smem;
  float* h = te_smem + d.dim;
  const int n = blockIdx.x;
  const int half = d.dim / 2;
  const float t = d.ts[n];
  for (int j = threadIdx.x; j < d.dim; j += blockDim.x) {
    const float a = __fmul_rn(t, d.freqs[j < half ? j : j - half]);
    e[j] = j < half ? cosf(a) : sinf(a);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < d.dim; o += nw) {
    const float* wr = d.w1 + (size_t)o * d.dim;
    float acc = 0.f;
    for (int j = lane; j < d.dim; j += 32) acc += wr[j] * e[j];
    acc = warp_sum(acc);
    if (lane == 0) h[o] = gelu_erf(acc + d.b1[o]);
  }
  __syncthreads();
  for (int o = warp; o < d.dim; o += nw) {
    const float* wr = d.w2 + (size_t)o * d.dim;
    float acc = 0.f;
    for (int j = lane; j < d.dim; j += 32) acc += wr[j] * h[j];
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + d.b2[o];
      if (d.class_embed) v += d.class_embed[(size_t)d.labels[n] * d.dim + o];
      d.emb[(size_t)n * d.dim + o] = v;
      d.gelu_emb[(size_t)n * d.dim + o] = gelu_erf(v);
    }
  }
}

__global__ void gelu_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = gelu_erf(in[i]);
}

// ab[n, o] = b[o] + sum_j w[o, j] * g[n, j]: a [n_out x dim] x [dim x batch] product (26 000 x 256 x 64 for unet64) as a
// register-tiled CUDA-core GEMM: 64 outputs x 64 samples per CTA, 4 x 4 per thread, K staged 16 at a time.  (One warp per
// output with a shuffle reduction per sample took 0.36 ms of every UNet step.)
constexpr int FL_TILE = 64, FL_K = 16;
__global__ void __launch_bounds__(256) film_linear_kernel(VqvsFilm d) {
  __shared__ float ws[FL_K][FL_TILE + 1], gs[FL_K][FL_TILE + 1];
  const int o0 = blockIdx.x * FL_TILE, n0 = blockIdx.y * FL_TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> 4 outputs, ty -> 4 samples
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d.dim; k0 += FL_K) {
    for (int i = threadIdx.x; i < FL_TILE * FL_K; i += 256) {
      const int r = i / FL_K, k = i - r * FL_K;  // consecutive threads read consecutive k of one row: 64-B segments
      ws[k][r] = (o0 + r < d.n_out && k0 + k < d.dim) ? d.w_cat[(size_t)(o0 + r) * d.dim + k0 + k] : 0.f;
      gs[k][r] = (n0 + r < d.batch && k0 + k < d.dim) ? d.gelu_emb[(size_t)(n0 + r) * d.dim + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FL_K; ++k) {
      float wv[4], gv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        wv[i] = ws[k][tx + 16 * i];
        gv[i] = gs[k][ty + 16 * i];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(gv[a], wv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = n0 + ty + 16 * a;
    if (n >= d.batch) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int o = o0 + tx + 16 * b;
      if (o < d.n_out) d.ab[(size_t)n * d.n_out + o] = acc[a][b] + d.b_cat[o];
    }
  }
}

// =============================================================================
// VQ: nearest codebook entry (warp-shuffle argmin) and embedding gather
// =============================================================================
constexpr int VQ_VEC = 4;      // vectors per CTA
constexpr int VQ_THREADS = 256;

__global__ void __launch_bounds__(VQ_THREADS) vq_argmin_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ dict, int n_vec,
                                                                int c, int t1, int dsize,
                                                                int64_t* __restrict__ idx) {
  extern __shared__ float vq_smem[];  // xs[VQ_VEC][c]
  __shared__ float xnorm[VQ_VEC];
  __shared__ float best_d[VQ_THREADS / 32][VQ_VEC];
  __shared__ int best_i[VQ_THREADS / 32][VQ_VEC];
  const int v0 = blockIdx.x * VQ_VEC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < VQ_VEC * c; i += VQ_THREADS) {
    const int v = i / c, ch = i - v * c;
    const int gv = v0 + v;
    float val = 0.f;
    if (gv < n_vec) {
      const int n = gv / t1, t = gv - n * t1;
      val = x[((size_t)n * c + ch) * t1 + t];
    }
    vq_smem[v * c + ch] = val;
  }
  __syncthreads();
  if (warp < VQ_VEC) {  // |x|^2 in fp64, rounded once (reference: torch.sum(torch.pow(x,2)), vq.py:213)
    double s = 0.0;
    for (int j = lane; j < c; j += 32) s += (double)vq_smem[warp * c + j] * (double)vq_smem[warp * c + j];
    s = warp_sum(s);
    if (lane == 0) xnorm[warp] = (float)s;
  }
  __syncthreads();
  float bd[VQ_VEC];
  int bi[VQ_VEC];
#pragma unroll
  for (int v = 0; v < VQ_VEC; ++v) { bd[v] = INFINITY; bi[v] = 0x7fffffff; }
  for (int row = warp; row < dsize; row += VQ_THREADS / 32) {
    const float* dr = dict + (size_t)row * c;
    double dot[VQ_VEC] = {0.0, 0.0, 0.0, 0.0};
    double dn = 0.0;
    for (int j = lane; j < c; j += 32) {
      const double w = (double)dr[j];
      dn += w * w;
#pragma unroll
      for (int v = 0; v < VQ_VEC; ++v) dot[v] += w * (double)vq_smem[v * c + j];
    }
    dn = warp_sum(dn);
#pragma unroll
    for (int v = 0; v < VQ_VEC; ++v) {
      const double dv = warp_sum(dot[v]);
      // ((-2*dots) + dict_norms) + tensor_norms, each step rounded to fp32 (vq.py:221)
      const float dist = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, (float)dv), (float)dn), xnorm[v]);
      if (dist < bd[v]) { bd[v] = dist; bi[v] = row; }  // rows ascend within a warp: first min kept
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v < VQ_VEC; ++v) { best_d[warp][v] = bd[v]; best_i[warp][v] = bi[v]; }
  }
  __syncthreads();
  if (threadIdx.x < VQ_VEC && v0 + threadIdx.x < n_vec) {
    const int v = threadIdx.x;
    float d0 = best_d[0][v];
    int i0 = best_i[0][v];
    for (int w = 1; w < VQ_THREADS / 32; ++w) {
      const float dw = best_d[w][v];
      const int iw = best_i[w][v];
      if (dw < d0 || (dw == d0 && iw < i0)) { d0 = dw; i0 = iw; }
    }
    idx[v0 + v] = (int64_t)i0;
  }
}

__global__ void vq_embed_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dict, int n, int c,
                                int t1, int d, float* __restrict__ out) {
  const size_t total = (size_t)n * c * t1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % t1);
    const size_t nc = i / t1;
    const int ch = (int)(nc % c);
    const int b = (int)(nc / c);
    // F.embedding raises on an out-of-range code; a kernel cannot, so it never reads outside the dictionary and marks the
    // element NaN (the Python wrapper checks the range up front and raises IndexError like the reference)
    const int64_t code = idx[(size_t)b * t1 + t];
    out[i] = (code >= 0 && code < d) ? dict[(size_t)code * c + ch] : __int_as_float(0x7fc00000);
  }
}

// =============================================================================
// keyed Gaussian noise (sharding-invariant sampling)
// =============================================================================
// out[r, j] ~ N(0, 1) as a pure function of (seed, first_row + r, step, j): Philox4x32-10 with key = seed and counter =
// (global row, step, j / 4); the four 32-bit words become two Box-Muller pairs.  A batch therefore draws the same noise
// however it is sharded over GPUs, with no host RNG, no H2D copy and no dependence on the device generator's state.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* o) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__global__ void keyed_normal_kernel(float* __restrict__ out, int rows, long long length, unsigned long long seed,
                                    long long first_row, int step) {
  const long long quads = (length + 3) / 4;
  const long long total = (long long)rows * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / quads);
    const long long q = i - (long long)r * quads;
    const unsigned long long row = (unsigned long long)(first_row + r);
    uint32_t w[4];
    philox4x32_10((uint32_t)row, (uint32_t)(row >> 32) ^ ((uint32_t)step * 0x85EBCA6Bu), (uint32_t)q, (uint32_t)(q >> 32) + (uint32_t)step,
                  (uint32_t)seed, (uint32_t)(seed >> 32), w);
    float z[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float u1 = ((float)(w[2 * p] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
      const float u2 = ((float)(w[2 * p + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincosf(6.283185307179586f * u2, &sn, &cs);
      z[2 * p] = rad * cs;
      z[2 * p + 1] = rad * sn;
    }
    float* dst = out + (long long)r * length + 4 * q;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (4 * q + e < length) dst[e] = z[e];
  }
}

// =============================================================================
// general fused conv (fp32 CUDA cores)
// =============================================================================
constexpr int SC_CO = 64;       // output channels per CTA
constexpr int SC_T = 128;       // output positions per CTA
constexpr int SC_CI = 8;        // input channels per smem stage
constexpr int SC_THREADS = 256;
constexpr int SC_MAXHALO = 64;  // 2 * max dilation
constexpr int SC_AW = SC_T + SC_MAXHALO;

// value of the (activated, resized) conv input at position tc of channel c, sample n
__device__ __forceinline__ float conv_src(const float* __restrict__ xa, const float* __restrict__ xb, int c_a,
                                          int c_b, int t_in, int n, int c, int tc, int t_conv, int resize,
                                          bool act, float sc, float sh) {
  if (tc < 0 || tc >= t_conv) return 0.f;
  const float* p = c < c_a ? xa + ((size_t)n * c_a + c) * t_in : xb + ((size_t)n * c_b + (c - c_a)) * t_in;
  if (resize == VQVS_RESIZE_NONE) {
    const float v = p[tc];
    return act ? gelu_erf(v * sc + sh) : v;
  } else if (resize == VQVS_RESIZE_UP2) {
    const float v = p[tc >> 1];
    return act ? gelu_erf(v * sc + sh) : v;
  } else {
    const float v0 = p[2 * tc], v1 = p[2 * tc + 1];
    if (act) return 0.5f * (gelu_erf(v0 * sc + sh) + gelu_erf(v1 * sc + sh));
    return 0.5f * (v0 + v1);
  }
}

__global__ void __launch_bounds__(SC_THREADS) conv1d_simt_kernel(VqvsConv d) {
  __shared__ float acts[SC_CI][SC_AW];
  __shared__ __align__(16) float ws[SC_CI][3][SC_CO];
  const int n = blockIdx.z;
  const int co0 = blockIdx.y * SC_CO;
  const int t0 = blockIdx.x * SC_T;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c_in = d.c_a + d.c_b;
  const int pad = (d.ksize / 2) * d.dilation;
  const int aw = SC_T + 2 * pad;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // ---- main taps over the activated input ------------------------------------
  for (int c0 = 0; c0 < c_in; c0 += SC_CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < SC_CI * aw; i += SC_THREADS) {
      const int ci = i / aw, col = i - ci * aw;
      const int c = c0 + ci;
      float v = 0.f;
      if (c < c_in) {
        const float sc = d.act ? d.scale[n * c_in + c] : 1.f;
        const float sh = d.act ? d.shift[n * c_in + c] : 0.f;
        v = conv_src(d.xa, d.xb, d.c_a, d.c_b, d.t_in, n, c, t0 - pad + col, d.t_out, d.resize, d.act != 0, sc, sh);
      }
      acts[ci][col] = v;
    }
    for (int i = threadIdx.x; i < SC_CI * d.ksize * SC_CO; i += SC_THREADS) {
      const int co = i % SC_CO;
      const int k = (i / SC_CO) % d.ksize;
      const int ci = i / (SC_CO * d.ksize);
      const int c = c0 + ci, o = co0 + co;
      ws[ci][k][co] = (c < c_in && o < d.c_out) ? d.w[((size_t)o * c_in + c) * d.ksize + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < SC_CI; ++ci) {
      for (int k = 0; k < d.ksize; ++k) {
        float a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = acts[ci][tx + 32 * j + k * d.dilation];
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[ci][k][ty * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[ci][k][ty * 8 + 4]);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(w[i], a[j], acc[i][j]);
      }
    }
  }

  // ---- 1x1 skip conv over the raw (resized) skip input -------------------------
  if (d.skip_mode == VQVS_SKIP_CONV1X1) {
    const int s_in = d.s_a + d.s_b;
    for (int c0 = 0; c0 < s_in; c0 += SC_CI) {
      __syncthreads();
      for (int i = threadIdx.x; i < SC_CI * SC_T; i += SC_THREADS) {
        const int ci = i / SC_T, col = i - ci * SC_T;
        const int c = c0 + ci;
        acts[ci][col] = c < s_in ? conv_src(d.sa, d.sb, d.s_a, d.s_b, d.t_skip, n, c, t0 + col, d.t_out,
                                            d.skip_resize, false, 1.f, 0.f)
                                 : 0.f;
      }
      for (int i = threadIdx.x; i < SC_CI * SC_CO; i += SC_THREADS) {
        const int co = i % SC_CO, ci = i / SC_CO;
        const int c = c0 + ci, o = co0 + co;
        ws[ci][0][co] = (c < s_in && o < d.c_out) ? d.w_skip[(size_t)o * s_in + c] : 0.f;
      }
      __syncthreads();
#pragma unroll 2
      for (int ci = 0; ci < SC_CI; ++ci) {
        float a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = acts[ci][tx + 32 * j];
        const float4 w0 = *reinterpret_cast<const float4*>(&ws[ci][0][ty * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&ws[ci][0][ty * 8 + 4]);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(w[i], a[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue: bias, identity skip, store, channel statistics ----------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = co0 + ty * 8 + i;
    if (o >= d.c_out) continue;  // warp-uniform
    float b = d.bias ? d.bias[o] : 0.f;
    if (d.skip_mode == VQVS_SKIP_CONV1X1 && d.b_skip) b += d.b_skip[o];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + 32 * j;
      if (t < d.t_out) {
        float v = acc[i][j] + b;
        if (d.skip_mode == VQVS_SKIP_IDENTITY)
          v += conv_src(d.sa, d.sb, d.s_a, d.s_b, d.t_skip, n, o, t, d.t_out, d.skip_resize, false, 1.f, 0.f);
        d.out[((size_t)n * d.c_out + o) * d.t_out + t] = v;
        s += v;
        ss += v * v;
      }
    }
    if (d.stats_out) {
      s = warp_sum(s);
      ss = warp_sum(ss);
      if (tx == 0) {
        atomicAdd(d.stats_out + ((size_t)n * d.c_out + o) * 2, (double)s);
        atomicAdd(d.stats_out + ((size_t)n * d.c_out + o) * 2 + 1, (double)ss);
      }
    }
  }
}

}  // namespace vqvs

// =============================================================================
// C ABI
// =============================================================================
using namespace vqvs;

static int check_conv(const VqvsConv* d) {
  VQVS_CHECK_ARG(d != nullptr, "conv: null descriptor");
  VQVS_CHECK_ARG(d->batch > 0 && d->c_a > 0 && d->c_b >= 0 && d->c_out > 0, "conv: bad channel/batch counts");
  VQVS_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv: ksize must be 1 or 3 (got %d)", d->ksize);
  VQVS_CHECK_ARG(d->dilation >= 1 && d->dilation <= 32, "conv: dilation %d outside [1,32]", d->dilation);
  const int expect = d->resize == VQVS_RESIZE_DOWN2 ? d->t_in / 2 : d->resize == VQVS_RESIZE_UP2 ? d->t_in * 2 : d->t_in;
  VQVS_CHECK_ARG(d->t_in > 0 && d->t_out == expect, "conv: t_out %d does not match t_in %d under resize %d", d->t_out,
                 d->t_in, d->resize);
  VQVS_CHECK_ARG(d->xa && (d->c_b == 0 || d->xb) && d->out, "conv: null tensor pointer");
  VQVS_CHECK_ARG(!d->act || (d->scale && d->shift), "conv: act=1 needs scale/shift");
  if (d->skip_mode != VQVS_SKIP_NONE) {
    VQVS_CHECK_ARG(d->sa && d->s_a > 0 && (d->s_b == 0 || d->sb), "conv: skip sources missing");
    const int sexp = d->skip_resize == VQVS_RESIZE_DOWN2 ? d->t_skip / 2 : d->skip_resize == VQVS_RESIZE_UP2 ? d->t_skip * 2 : d->t_skip;
    VQVS_CHECK_ARG(d->t_skip > 0 && sexp == d->t_out, "conv: skip length %d does not give t_out %d under resize %d", d->t_skip, d->t_out, d->skip_resize);
    if (d->skip_mode == VQVS_SKIP_IDENTITY)
      VQVS_CHECK_ARG(d->s_a + d->s_b == d->c_out, "conv: identity skip needs %d channels, got %d", d->c_out, d->s_a + d->s_b);
  }
  return VQVS_OK;
}

extern "C" int vqvs_conv1d_fused(const VqvsConv* d, void* stream) {
  int rc = check_conv(d);
  if (rc) return rc;
  VQVS_CHECK_ARG(d->w, "conv(simt): fp32 weights missing");
  VQVS_CHECK_ARG(d->skip_mode != VQVS_SKIP_CONV1X1 || d->w_skip, "conv(simt): fp32 skip weights missing");
  dim3 grid(ceil_div(d->t_out, SC_T), ceil_div(d->c_out, SC_CO), d->batch);
  conv1d_simt_kernel<<<grid, SC_THREADS, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_conv1d_fused");
  return VQVS_OK;
}

extern "C" int vqvs_gn_finalize(const VqvsGnFinalize* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c_a > 0 && d->c_b >= 0, "gn_finalize: bad sizes");
  const int C = d->c_a + d->c_b;
  VQVS_CHECK_ARG(d->groups > 0 && C % d->groups == 0, "gn_finalize: %d channels not divisible into %d groups", C, d->groups);
  VQVS_CHECK_ARG(d->stats_a && (d->c_b == 0 || d->stats_b) && d->gamma && d->beta && d->scale && d->shift,
                 "gn_finalize: null pointer");
  VQVS_CHECK_ARG(d->count > 0, "gn_finalize: count must be positive");
  const int total = d->batch * C;
  gn_finalize_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_gn_finalize");
  return VQVS_OK;
}

extern "C" int vqvs_channel_stats(const float* x, int batch, int c, int t, double* stats, void* stream) {
  VQVS_CHECK_ARG(x && stats && batch > 0 && c > 0 && t > 0, "channel_stats: bad arguments");
  channel_stats_kernel<<<batch * c, 256, 0, (cudaStream_t)stream>>>(x, t, stats);
  VQVS_CHECK_LAUNCH("vqvs_channel_stats");
  return VQVS_OK;
}

extern "C" int vqvs_conv_in(const VqvsConvIn* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c_out > 0 && d->t > 0, "conv_in: bad sizes");
  VQVS_CHECK_ARG(d->x && d->w && d->bias && d->out, "conv_in: null pointer");
  VQVS_CHECK_ARG(!d->cond || d->t_cond > 0, "conv_in: cond given without t_cond");
  const size_t smem = (size_t)(CIN_THREADS / 32) * d->c_out * 2 * sizeof(float);
  VQVS_CHECK_ARG(smem <= 48 * 1024, "conv_in: c_out %d too large", d->c_out);
  dim3 grid(ceil_div(d->t, CIN_THREADS * CIN_PER_THREAD), d->batch);
  conv_in_kernel<<<grid, CIN_THREADS, smem, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_conv_in");
  return VQVS_OK;
}

extern "C" int vqvs_conv_out(const VqvsConvOut* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c_in > 0 && d->t > 0, "conv_out: bad sizes");
  VQVS_CHECK_ARG(d->x && d->scale && d->shift && d->w && d->bias && d->out, "conv_out: null pointer");
  VQVS_CHECK_ARG(d->mode >= VQVS_OUT_EPS && d->mode <= VQVS_OUT_X0_SUM, "conv_out: bad mode %d", d->mode);
  if (d->mode != VQVS_OUT_EPS) VQVS_CHECK_ARG(d->x_t && d->coef, "conv_out: mode %d needs x_t and coef", d->mode);
  if (d->mode == VQVS_OUT_X0_SUM) VQVS_CHECK_ARG(d->x0_sum, "conv_out: x0_sum missing");
  if (d->t % COUT_VEC == 0 && (reinterpret_cast<uintptr_t>(d->x) & 15) == 0) {
    dim3 grid(ceil_div(d->t, COUT_THREADS * COUT_VEC), d->batch);
    conv_out_vec_kernel<<<grid, COUT_THREADS, 0, (cudaStream_t)stream>>>(*d);
  } else {
    dim3 grid(ceil_div(d->t, COUT_THREADS), d->batch);
    conv_out_kernel<<<grid, COUT_THREADS, 0, (cudaStream_t)stream>>>(*d);
  }
  VQVS_CHECK_LAUNCH("vqvs_conv_out");
  return VQVS_OK;
}

extern "C" int vqvs_ddpm_finish(const VqvsDdpmFinish* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->t > 0 && d->x_t && d->eps && d->coef && d->out, "ddpm_finish: bad arguments");
  VQVS_CHECK_ARG(!d->use_x0_mean || d->x0_sum, "ddpm_finish: x0_sum missing");
  dim3 grid(ceil_div(d->t, 256 * 4), d->batch);
  ddpm_finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_ddpm_finish");
  return VQVS_OK;
}

extern "C" int vqvs_ddpm_x0_sum(const float* x_t, const float* eps, const float* coef, int batch, int t,
                                double* x0_sum, void* stream) {
  VQVS_CHECK_ARG(x_t && eps && coef && x0_sum && batch > 0 && t > 0, "ddpm_x0_sum: bad arguments");
  dim3 grid(ceil_div(t, 256 * 8), batch);
  ddpm_x0_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_t, eps, coef, t, x0_sum);
  VQVS_CHECK_LAUNCH("vqvs_ddpm_x0_sum");
  return VQVS_OK;
}

extern "C" int vqvs_time_embed(const VqvsTimeEmbed* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->dim > 0 && d->dim % 2 == 0 && d->dim <= 4096, "time_embed: bad sizes");
  VQVS_CHECK_ARG(d->ts && d->freqs && d->w1 && d->b1 && d->w2 && d->b2 && d->emb && d->gelu_emb, "time_embed: null pointer");
  VQVS_CHECK_ARG((d->class_embed == nullptr) == (d->labels == nullptr), "time_embed: class_embed and labels go together");
  time_embed_kernel<<<d->batch, 256, 2 * d->dim * sizeof(float), (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_time_embed");
  return VQVS_OK;
}

extern "C" int vqvs_gelu(const float* in, float* out, int64_t n, void* stream) {
  VQVS_CHECK_ARG(in && out && n > 0, "gelu: bad arguments");
  const int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  gelu_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  VQVS_CHECK_LAUNCH("vqvs_gelu");
  return VQVS_OK;
}

static int launch_film(const VqvsFilm& f, void* stream) {
  VQVS_CHECK_ARG(f.gelu_emb && f.w_cat && f.b_cat && f.ab && f.batch > 0 && f.n_out > 0, "film_linear: bad arguments");
  VQVS_CHECK_ARG(f.dim > 0, "film_linear: dim must be positive");
  film_linear_kernel<<<dim3(ceil_div(f.n_out, FL_TILE), ceil_div(f.batch, FL_TILE)), 256, 0, (cudaStream_t)stream>>>(f);
  VQVS_CHECK_LAUNCH("vqvs_film_linear");
  return VQVS_OK;
}

extern "C" int vqvs_film_linear(const float* gelu_emb, const float* w_cat, const float* b_cat, int batch, int dim,
                                int n_out, float* ab, void* stream) {
  VqvsFilm f{gelu_emb, w_cat, b_cat, batch, dim, n_out, ab};
  return launch_film(f, stream);
}

extern "C" int vqvs_vq_argmin(const float* x, const float* dict, int n, int c, int t1, int d, int64_t* idx,
                              void* stream) {
  VQVS_CHECK_ARG(x && dict && idx && n > 0 && c > 0 && t1 > 0 && d > 0, "vq_argmin: bad arguments");
  const size_t smem = (size_t)VQ_VEC * c * sizeof(float);
  VQVS_CHECK_ARG(smem <= 48 * 1024, "vq_argmin: %d channels too many", c);
  const int n_vec = n * t1;
  vq_argmin_kernel<<<ceil_div(n_vec, VQ_VEC), VQ_THREADS, smem, (cudaStream_t)stream>>>(x, dict, n_vec, c, t1, d, idx);
  VQVS_CHECK_LAUNCH("vqvs_vq_argmin");
  return VQVS_OK;
}

extern "C" int vqvs_vq_embed(const int64_t* idx, const float* dict, int n, int c, int t1, int d, float* out,
                             void* stream) {
  VQVS_CHECK_ARG(idx && dict && out && n > 0 && c > 0 && t1 > 0 && d > 0, "vq_embed: bad arguments");
  const size_t total = (size_t)n * c * t1;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  vq_embed_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, dict, n, c, t1, d, out);
  VQVS_CHECK_LAUNCH("vqvs_vq_embed");
  return VQVS_OK;
}

extern "C" int vqvs_keyed_normal(float* out, int rows, int64_t length, uint64_t seed, int64_t first_row, int32_t step,
                                 void* stream) {
  VQVS_CHECK_ARG(rows >= 0 && length >= 0 && (out || rows == 0 || length == 0), "keyed_normal: bad arguments");
  if (rows == 0 || length == 0) return VQVS_OK;
  const long long total = (long long)rows * ((length + 3) / 4);
  const int blocks = (int)(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  keyed_normal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, rows, (long long)length, (unsigned long long)seed,
                                                               (long long)first_row, (int)step);
  VQVS_CHECK_LAUNCH("vqvs_keyed_normal");
  return VQVS_OK;
}

namespace vqvs {
int run_film(const VqvsFilm* f, void* stream) { return launch_film(*f, stream); }
}  // namespace vqvs
