// Classifier guidance (BASELINE configs[4]): the kernels that take the guidance model of reference
// models/classifier.py off ATen/autograd.  The convolutions of the stem -- forward AND the input-gradient ("dgrad")
// ones -- run on vqvs_conv1d_umma (a transposed conv is the same implicit GEMM with flipped taps and transposed packed
// weights); this file holds what sits between them on the way back:
//
//   forward  y = skip(x) + conv2(gelu(film(GN_b(conv1(resize(gelu(GN_a(x))))))))          (models/unet.py:307-316)
//   backward dw = conv2^T(dy) -> [gelu' , FiLM, GN_b]^T -> du -> dp = conv1^T(du) -> resize^T -> [gelu', GN_a]^T -> dx
//
// GroupNorm's backward needs two per-group reductions of the incoming gradient, so it splits exactly like the forward:
//   vqvs_gelu_bwd   q = d * gelu'(z*S + H) * gamma(1+a), and per (n, c): sum_t q, sum_t q*zhat      (producer side)
//   vqvs_gn_bwd_finalize   per (n, c) coefficients (A, B, C) of  dz = A*q + B*z + C                 (tiny)
//   vqvs_affine3    dz = A*q + B*z + C (+ the skip path's gradient, through resize^T)               (consumer side)
// plus the attention pool (classifier.py:133-191; only query row 0 exists and it is the bias because token 0 is the zero
// pad), the Linear head and the 1 -> C input conv.  All of these are HBM- or latency-bound fp32 CUDA-core kernels.
#include "common.cuh"

namespace vqvs {

__device__ __forceinline__ float gelu_grad(float v) {  // d/dv [v * Phi(v)] = Phi(v) + v * phi(v)
  // Phi(-|v|) = 0.5 erfc(|v| / sqrt2) in the Abramowitz-Stegun 7.1.26 form (|error| <= 1e-7; Phi itself, not v * Phi, is
  // needed here, which is what the prologues' 2^P5 fit is NOT tuned for near 0); its
  // exp(-v^2 / 2) is also the density's, so one ex2 + one rcp serve both terms (erff + expf cost ~3x the instructions).
  const float av = fabsf(v);
  const float t = rcp_approx(fmaf(0.2316418882f, av, 1.0f));
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float e = ex2_approx((v * v) * -0.72134752044f);  // exp(-v^2 / 2)
  const float h = (t * p) * e;                             // Phi(-|v|)
  const float cdf = v >= 0.f ? 1.0f - h : h;
  return fmaf(v, 0.3989422804014327f * e, cdf);
}

__device__ __forceinline__ double block_sum(double v, double* red) {  // blockDim.x <= 1024; result valid in thread 0
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (int)((blockDim.x + 31) >> 5) ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// ---- per-(n, c) quantities of one GroupNorm(+FiLM): prep = [S | H | mean | rstd | gf], each [batch * C] -------------
__global__ void gn_bwd_prep_kernel(VqvsGnFinalize d, float* __restrict__ prep) {
  const int C = d.c_a + d.c_b;
  const int total = d.batch * C;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = idx / C, c = idx - n * C;
  const int cg = C / d.groups, g = c / cg;
  double s = 0.0, ss = 0.0;
  for (int cc = g * cg; cc < (g + 1) * cg; ++cc) {
    const double* st = cc < d.c_a ? d.stats_a + ((size_t)n * d.c_a + cc) * 2 : d.stats_b + ((size_t)n * d.c_b + (cc - d.c_a)) * 2;
    s += st[0];
    ss += st[1];
  }
  const double cnt = (double)cg * (double)d.count;
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double rstd = rsqrt(var + 1e-5);
  double gf = (double)d.gamma[c];
  double sc = rstd * gf, sh = (double)d.beta[c] - mean * sc;
  if (d.film) {
    const double a = (double)d.film[(size_t)n * d.film_stride + c], b = (double)d.film[(size_t)n * d.film_stride + C + c];
    sc *= 1.0 + a;
    sh = sh * (1.0 + a) + b;
    gf *= 1.0 + a;
  }
  prep[idx] = (float)sc;
  prep[total + idx] = (float)sh;
  prep[2 * total + idx] = (float)mean;
  prep[3 * total + idx] = (float)rstd;
  prep[4 * total + idx] = (float)gf;
}

// ---- q = d * gelu'(z*S + H) * gf ; acc[row] += (sum q, sum q*zhat) -------------------------------------------------------
// z / q are dense [batch, c, t]; the GroupNorm may span a channel CONCATENATION of c_total channels of which this source is
// [c_off, c_off + c): prep / acc and the incoming gradient d_in [batch, c_total, t_d] are indexed through (c_total, c_off).
// up: 0 = d_in has length t; 1 = the forward avg-pooled (d_in length t/2, reaches i as 0.5*d_in[i/2]);
//     2 = the forward upsampled x2 (d_in length 2t, position i collects d_in[2i] + d_in[2i+1]).
constexpr int GB_THREADS = 256, GB_CHUNK = 4096;
__global__ void __launch_bounds__(GB_THREADS) gelu_bwd_kernel(VqvsGeluBwd d) {
  __shared__ double red[32];
  const int row = blockIdx.y, total = d.batch * d.c_total;
  const int n = row / d.c, prow = n * d.c_total + d.c_off + (row - n * d.c);
  const float S = d.prep[prow], H = d.prep[total + prow], mean = d.prep[2 * total + prow], rstd = d.prep[3 * total + prow],
              gf = d.prep[4 * total + prow];
  const float* z = d.z + (size_t)row * d.t;
  const float* din = d.d_in + (size_t)prow * (d.up == 1 ? d.t / 2 : d.up == 2 ? 2 * d.t : d.t);
  float* q = d.q + (size_t)row * d.t;
  const int t0 = blockIdx.x * GB_CHUNK, t1 = min(d.t, t0 + GB_CHUNK);
  float s1 = 0.f, s2 = 0.f;
  // 16-byte path: un-resampled gradient, rows 16-byte aligned (these kernels are HBM-bound: 12 B per element)
  const bool vec = d.up == 0 && (d.t & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.z) | reinterpret_cast<uintptr_t>(d.d_in) |
                                                   reinterpret_cast<uintptr_t>(d.q)) & 15) == 0;
  if (vec) {
    for (int i = t0 + 4 * threadIdx.x; i < t1; i += 4 * GB_THREADS) {
      const float4 z4 = *reinterpret_cast<const float4*>(z + i), g4 = *reinterpret_cast<const float4*>(din + i);
      const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] = gg[e] * gelu_grad(fmaf(zz[e], S, H)) * gf;
        s1 += v[e];
        s2 = fmaf(v[e], (zz[e] - mean) * rstd, s2);
      }
      *reinterpret_cast<float4*>(q + i) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (int i = t0 + threadIdx.x; i < t1; i += GB_THREADS) {
      const float zz = z[i];
      float g;
      if (d.up == 1) g = 0.5f * din[i >> 1];
      else if (d.up == 2) {
        const float2 pr = *reinterpret_cast<const float2*>(din + 2 * i);
        g = pr.x + pr.y;
      } else g = din[i];
      const float v = g * gelu_grad(fmaf(zz, S, H)) * gf;
      q[i] = v;
      s1 += v;
      s2 = fmaf(v, (zz - mean) * rstd, s2);
    }
  }
  const double r1 = block_sum((double)s1, red), r2 = block_sum((double)s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(d.acc + (size_t)prow * 2, r1);
    atomicAdd(d.acc + (size_t)prow * 2 + 1, r2);
  }
}

// ---- (A, B, C) of dz = A*q + B*z + C = rstd*(q - m1 - zhat*m2), m1/m2 the group means of q and q*zhat ---------------------
__global__ void gn_bwd_finalize_kernel(VqvsGnBwdFinalize d) {
  const int total = d.batch * d.c;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = idx / d.c, c = idx - n * d.c;
  const int cg = d.c / d.groups, g = c / cg;
  double a1 = 0.0, a2 = 0.0;
  for (int cc = g * cg; cc < (g + 1) * cg; ++cc) {
    a1 += d.acc[((size_t)n * d.c + cc) * 2];
    a2 += d.acc[((size_t)n * d.c + cc) * 2 + 1];
  }
  const double cnt = (double)cg * (double)d.count;
  const double m1 = a1 / cnt, m2 = a2 / cnt;
  const double mean = d.prep[2 * total + idx], rstd = d.prep[3 * total + idx];
  d.coef[idx] = (float)rstd;
  d.coef[total + idx] = (float)(-rstd * rstd * m2);
  d.coef[2 * total + idx] = (float)(-rstd * m1 + rstd * rstd * m2 * mean);
}

// ---- out = A*q + B*z + C (+ add through resize^T) (+ add2) ----------------------------------------------------------------
// coef and add [batch, c_total, t_add] are indexed through (c_total, c_off) like vqvs_gelu_bwd's d_in; add2 is dense.
__global__ void __launch_bounds__(GB_THREADS) affine3_kernel(VqvsAffine3 d) {
  const int row = blockIdx.y, total = d.batch * d.c_total;
  const int n = row / d.c, prow = n * d.c_total + d.c_off + (row - n * d.c);
  const float A = d.coef[prow], B = d.coef[total + prow], Cc = d.coef[2 * total + prow];
  const float* q = d.q + (size_t)row * d.t;
  const float* z = d.z + (size_t)row * d.t;
  const float* add = d.add ? d.add + (size_t)prow * (d.add_mode == 2 ? d.t / 2 : d.add_mode == 3 ? 2 * d.t : d.t) : nullptr;
  const float* add2 = d.add2 ? d.add2 + (size_t)row * d.t : nullptr;
  float* out = d.out + (size_t)row * d.t;
  const int t0 = blockIdx.x * GB_CHUNK, t1 = min(d.t, t0 + GB_CHUNK);
  const bool vec = d.add_mode <= 1 && (d.t & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(d.q) | reinterpret_cast<uintptr_t>(d.z) | reinterpret_cast<uintptr_t>(d.out) |
                     reinterpret_cast<uintptr_t>(d.add) | reinterpret_cast<uintptr_t>(d.add2)) & 15) == 0;
  if (vec) {  // 16-byte path (12-20 B per element, HBM-bound)
    for (int i = t0 + 4 * threadIdx.x; i < t1; i += 4 * GB_THREADS) {
      const float4 q4 = *reinterpret_cast<const float4*>(q + i), z4 = *reinterpret_cast<const float4*>(z + i);
      float4 o = make_float4(fmaf(A, q4.x, fmaf(B, z4.x, Cc)), fmaf(A, q4.y, fmaf(B, z4.y, Cc)), fmaf(A, q4.z, fmaf(B, z4.z, Cc)),
                             fmaf(A, q4.w, fmaf(B, z4.w, Cc)));
      if (d.add_mode == 1) {
        const float4 a = *reinterpret_cast<const float4*>(add + i);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      if (add2) {
        const float4 a = *reinterpret_cast<const float4*>(add2 + i);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      *reinterpret_cast<float4*>(out + i) = o;
    }
    return;
  }
  for (int i = t0 + threadIdx.x; i < t1; i += GB_THREADS) {
    float v = fmaf(A, q[i], fmaf(B, z[i], Cc));
    if (d.add_mode == 1) v += add[i];
    else if (d.add_mode == 2) v += 0.5f * add[i >> 1];
    else if (d.add_mode == 3) v += add[2 * i] + add[2 * i + 1];
    if (add2) v += add2[i];
    out[i] = v;
  }
}

// ---- input conv backward: dx[n, t] = sum_c sum_k w[c, 0, k] * dh[n, c, t + 1 - k]   (Conv1d(1 -> C, k = 3, pad 1)) ----------
__global__ void conv_in_bwd_kernel(VqvsConvInBwd d) {
  extern __shared__ float w_s[];  // [c][3]
  for (int i = threadIdx.x; i < d.c * 3; i += blockDim.x) w_s[i] = d.w[i];
  __syncthreads();
  const int n = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.t) return;
  const float* dh = d.dh + (size_t)n * d.c * d.t;
  float acc = 0.f;
  for (int c = 0; c < d.c; ++c) {
    const float* r = dh + (size_t)c * d.t;
    const float l = t + 1 < d.t ? r[t + 1] : 0.f, m = r[t], u = t > 0 ? r[t - 1] : 0.f;
    acc = fmaf(w_s[c * 3], l, fmaf(w_s[c * 3 + 1], m, fmaf(w_s[c * 3 + 2], u, acc)));
  }
  d.dx[(size_t)n * d.t + t] = acc;
}

// ---- attention pool (classifier.py:133-191) ----------------------------------------------------------------------------------
// tokens X[c, s]: X[:, 0] = 0, X[c, s] = gelu(h[c, s-1]*S[c] + H[c]).  Only query row 0 is consumed and X[:, 0] = 0, so
// q0 = b_q; with qk[h, c] = sum_j q0[h, j] Wk[h*ch + j, c]:
//   logit[h, s] = scale2 * (qk[h, :].X[:, s] + q0[h, :].b_k[h, :]),  w = softmax_s,  xbar[h, c] = sum_s w[h, s] X[c, s]
//   pooled[h*ch + j] = Wv[h*ch + j, :].xbar[h, :] + b_v[h*ch + j],   out = Wc.pooled + b_c
// One CTA per sample; ws (per sample): w [heads][t+1] | xbar [heads][c] | pooled [c] | qk [heads][c]
constexpr int AP_THREADS = 256;
__device__ __forceinline__ float ap_token(const VqvsAttnPool& d, int n, int c, int s, const float* S, const float* H) {
  if (s == 0) return 0.f;
  const float v = fmaf(d.h[((size_t)n * d.c + c) * d.t + (s - 1)], S[c], H[c]);
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}
__device__ __forceinline__ size_t ap_ws_floats(int c, int t, int heads) { return (size_t)heads * (t + 1) + 2 * (size_t)heads * c + c; }

__global__ void __launch_bounds__(AP_THREADS) attnpool_fwd_kernel(VqvsAttnPool d) {
  extern __shared__ float sm[];  // S[c] | H[c] | scratch[max(c, heads*(t+1))]
  const int n = blockIdx.x, C = d.c, Tt = d.t + 1, heads = d.heads, ch = C / heads, total = d.batch * C;
  float* S = sm;
  float* H = sm + C;
  float* ws = d.ws + (size_t)n * ap_ws_floats(C, d.t, heads);
  float* w = ws;                            // [heads][Tt]
  float* xbar = w + (size_t)heads * Tt;     // [heads][C]
  float* pooled = xbar + (size_t)heads * C; // [C]
  float* qk = pooled + C;                   // [heads][C]
  const float scale2 = rsqrtf((float)ch);   // (ch^-1/4)^2
  for (int c = threadIdx.x; c < C; c += AP_THREADS) {
    S[c] = d.prep[(size_t)n * C + c];
    H[c] = d.prep[total + (size_t)n * C + c];
  }
  for (int i = threadIdx.x; i < heads * C; i += AP_THREADS) {
    const int h = i / C, c = i - h * C;
    float a = 0.f;
    for (int j = 0; j < ch; ++j) a = fmaf(d.b_qkv[h * ch + j], d.w_qkv[(size_t)(C + h * ch + j) * C + c], a);
    qk[i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < heads * Tt; i += AP_THREADS) {
    const int h = i / Tt, s = i - h * Tt;
    float a = 0.f;
    for (int j = 0; j < ch; ++j) a = fmaf(d.b_qkv[h * ch + j], d.b_qkv[C + h * ch + j], a);
    if (s > 0)
      for (int c = 0; c < C; ++c) a = fmaf(qk[h * C + c], ap_token(d, n, c, s, S, H), a);
    w[i] = a * scale2;
  }
  __syncthreads();
  if (threadIdx.x < heads) {  // softmax over the t+1 keys of one head (126 entries)
    float* row = w + (size_t)threadIdx.x * Tt;
    float mx = row[0];
    for (int s = 1; s < Tt; ++s) mx = fmaxf(mx, row[s]);
    float sum = 0.f;
    for (int s = 0; s < Tt; ++s) {
      row[s] = expf(row[s] - mx);
      sum += row[s];
    }
    const float inv = 1.0f / sum;
    for (int s = 0; s < Tt; ++s) row[s] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < heads * C; i += AP_THREADS) {
    const int h = i / C, c = i - h * C;
    float a = 0.f;
    for (int s = 1; s < Tt; ++s) a = fmaf(w[h * Tt + s], ap_token(d, n, c, s, S, H), a);
    xbar[i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += AP_THREADS) {
    const int h = i / ch;
    float a = d.b_qkv[2 * C + i];
    const float* wv = d.w_qkv + (size_t)(2 * C + i) * C;
    for (int c = 0; c < C; ++c) a = fmaf(wv[c], xbar[h * C + c], a);
    pooled[i] = a;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < d.c_out; o += AP_THREADS) {
    float a = d.b_proj[o];
    const float* wc = d.w_proj + (size_t)o * C;
    for (int c = 0; c < C; ++c) a = fmaf(wc[c], pooled[c], a);
    d.out[(size_t)n * d.c_out + o] = a;
  }
}

// d_act[n, c, s-1] = dX[c, s] for s >= 1 (gradient w.r.t. the tokens gelu(GN(h)))
__global__ void __launch_bounds__(AP_THREADS) attnpool_bwd_kernel(VqvsAttnPool d) {
  extern __shared__ float sm[];  // S[c] | H[c] | dpooled[c] | dxbar[heads*c] | dlogit[heads*(t+1)]
  const int n = blockIdx.x, C = d.c, Tt = d.t + 1, heads = d.heads, ch = C / heads, total = d.batch * C;
  float* S = sm;
  float* H = S + C;
  float* dpooled = H + C;
  float* dxbar = dpooled + C;
  float* dlogit = dxbar + (size_t)heads * C;
  const float* ws = d.ws + (size_t)n * ap_ws_floats(C, d.t, heads);
  const float* w = ws;
  const float* qk = ws + (size_t)heads * Tt + (size_t)heads * C + C;
  const float scale2 = rsqrtf((float)ch);
  for (int c = threadIdx.x; c < C; c += AP_THREADS) {
    S[c] = d.prep[(size_t)n * C + c];
    H[c] = d.prep[total + (size_t)n * C + c];
    float a = 0.f;
    for (int o = 0; o < d.c_out; ++o) a = fmaf(d.w_proj[(size_t)o * C + c], d.d_out[(size_t)n * d.c_out + o], a);
    dpooled[c] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < heads * C; i += AP_THREADS) {
    const int h = i / C, c = i - h * C;
    float a = 0.f;
    for (int j = 0; j < ch; ++j) a = fmaf(d.w_qkv[(size_t)(2 * C + h * ch + j) * C + c], dpooled[h * ch + j], a);
    dxbar[i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < heads * Tt; i += AP_THREADS) {  // dw[h, s] = dxbar[h, :].X[:, s]
    const int h = i / Tt, s = i - h * Tt;
    float a = 0.f;
    if (s > 0)
      for (int c = 0; c < C; ++c) a = fmaf(dxbar[h * C + c], ap_token(d, n, c, s, S, H), a);
    dlogit[i] = a;
  }
  __syncthreads();
  if (threadIdx.x < heads) {  // softmax backward: dlogit = w * (dw - sum_s w*dw)
    const float* wr = w + (size_t)threadIdx.x * Tt;
    float* dr = dlogit + (size_t)threadIdx.x * Tt;
    float dot = 0.f;
    for (int s = 0; s < Tt; ++s) dot = fmaf(wr[s], dr[s], dot);
    for (int s = 0; s < Tt; ++s) dr[s] = wr[s] * (dr[s] - dot) * scale2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * d.t; i += AP_THREADS) {
    const int c = i / d.t, s = i - c * d.t + 1;
    float a = 0.f;
    for (int h = 0; h < heads; ++h) a = fmaf(w[h * Tt + s], dxbar[h * C + c], fmaf(dlogit[h * Tt + s], qk[h * C + c], a));
    d.d_act[((size_t)n * C + c) * d.t + (s - 1)] = a;
  }
}

// ---- head (classifier.py:36-45): logits = W gelu(stem) + b ; d_stem = gelu'(stem) * W^T d_logits ------------------------------
__global__ void cls_head_fwd_kernel(VqvsClsHead d) {
  const int n = blockIdx.x;
  for (int l = threadIdx.x; l < d.labels; l += blockDim.x) {
    float a = d.b[l];
    for (int i = 0; i < d.dim; ++i) {
      const float v = d.stem[(size_t)n * d.dim + i];
      a = fmaf(d.w[(size_t)l * d.dim + i], 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)), a);
    }
    d.logits[(size_t)n * d.labels + l] = a;
  }
}
__global__ void cls_head_bwd_kernel(VqvsClsHead d) {
  const int n = blockIdx.x;
  for (int i = threadIdx.x; i < d.dim; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < d.labels; ++l) a = fmaf(d.w[(size_t)l * d.dim + i], d.d_logits[(size_t)n * d.labels + l], a);
    d.d_stem[(size_t)n * d.dim + i] = a * gelu_grad(d.stem[(size_t)n * d.dim + i]);
  }
}

// ---- backward of h[:, :, j*rate] sampling (F.interpolate(h, size = t / rate), nearest): dh[j*rate] = d[j], zero elsewhere ----
__global__ void scatter_stride_kernel(const float* __restrict__ d, float* __restrict__ out, int t, int rate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = blockIdx.y;
  if (i >= t) return;
  out[row * t + i] = (i % rate == 0) ? d[row * (t / rate) + i / rate] : 0.f;
}
__global__ void gather_stride_kernel(const float* __restrict__ h, float* __restrict__ out, int t, int rate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = blockIdx.y;
  if (j >= t / rate) return;
  out[row * (t / rate) + j] = h[row * t + (size_t)j * rate];
}

}  // namespace vqvs

using namespace vqvs;

extern "C" int vqvs_stride_sample(const float* src, float* dst, int rows, int t, int rate, int backward, void* stream) {
  VQVS_CHECK_ARG(src && dst && rows > 0 && rows <= 65535 && t > 0 && rate > 0 && t % rate == 0, "stride_sample: bad arguments");
  if (backward) scatter_stride_kernel<<<dim3(ceil_div(t, 256), rows), 256, 0, (cudaStream_t)stream>>>(src, dst, t, rate);
  else gather_stride_kernel<<<dim3(ceil_div(t / rate, 256), rows), 256, 0, (cudaStream_t)stream>>>(src, dst, t, rate);
  VQVS_CHECK_LAUNCH("vqvs_stride_sample");
  return VQVS_OK;
}

extern "C" int vqvs_gn_bwd_prep(const VqvsGnFinalize* d, float* prep, void* stream) {
  VQVS_CHECK_ARG(d && prep && d->batch > 0 && d->groups > 0 && (d->c_a + d->c_b) % d->groups == 0 && d->count > 0,
                 "gn_bwd_prep: bad arguments");
  VQVS_CHECK_ARG(d->stats_a && (d->c_b == 0 || d->stats_b) && d->gamma && d->beta, "gn_bwd_prep: null pointer");
  const int total = d->batch * (d->c_a + d->c_b);
  gn_bwd_prep_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(*d, prep);
  VQVS_CHECK_LAUNCH("vqvs_gn_bwd_prep");
  return VQVS_OK;
}

extern "C" int vqvs_gelu_bwd(const VqvsGeluBwd* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c > 0 && d->t > 0 && (long long)d->batch * d->c <= 65535, "gelu_bwd: bad sizes");
  VQVS_CHECK_ARG(d->c_total >= d->c && d->c_off >= 0 && d->c_off + d->c <= d->c_total, "gelu_bwd: bad channel window");
  VQVS_CHECK_ARG(d->d_in && d->z && d->prep && d->q && d->acc, "gelu_bwd: null pointer");
  VQVS_CHECK_ARG(d->up >= 0 && d->up <= 2 && (d->up != 1 || (d->t & 1) == 0), "gelu_bwd: bad resize mode / odd pooled length");
  dim3 grid(ceil_div(d->t, GB_CHUNK), d->batch * d->c);
  gelu_bwd_kernel<<<grid, GB_THREADS, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_gelu_bwd");
  return VQVS_OK;
}

extern "C" int vqvs_gn_bwd_finalize(const VqvsGnBwdFinalize* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c > 0 && d->groups > 0 && d->c % d->groups == 0 && d->count > 0, "gn_bwd_finalize: bad sizes");
  VQVS_CHECK_ARG(d->acc && d->prep && d->coef, "gn_bwd_finalize: null pointer");
  gn_bwd_finalize_kernel<<<ceil_div(d->batch * d->c, 128), 128, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_gn_bwd_finalize");
  return VQVS_OK;
}

extern "C" int vqvs_affine3(const VqvsAffine3* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c > 0 && d->t > 0 && (long long)d->batch * d->c <= 65535, "affine3: bad sizes");
  VQVS_CHECK_ARG(d->q && d->z && d->coef && d->out && d->add_mode >= 0 && d->add_mode <= 3 && (d->add_mode == 0 || d->add),
                 "affine3: bad arguments");
  VQVS_CHECK_ARG(d->c_total >= d->c && d->c_off >= 0 && d->c_off + d->c <= d->c_total, "affine3: bad channel window");
  VQVS_CHECK_ARG(d->add_mode != 2 || (d->t & 1) == 0, "affine3: pooled skip gradient needs an even length");
  dim3 grid(ceil_div(d->t, GB_CHUNK), d->batch * d->c);
  affine3_kernel<<<grid, GB_THREADS, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_affine3");
  return VQVS_OK;
}

extern "C" int vqvs_conv_in_bwd(const VqvsConvInBwd* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c > 0 && d->t > 0 && d->dh && d->w && d->dx, "conv_in_bwd: bad arguments");
  VQVS_CHECK_ARG((size_t)d->c * 3 * sizeof(float) <= 48 * 1024, "conv_in_bwd: too many channels");
  dim3 grid(ceil_div(d->t, 256), d->batch);
  conv_in_bwd_kernel<<<grid, 256, d->c * 3 * sizeof(float), (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_conv_in_bwd");
  return VQVS_OK;
}

extern "C" int64_t vqvs_attnpool_workspace_bytes(int batch, int c, int t, int heads) {
  if (batch <= 0 || c <= 0 || t <= 0 || heads <= 0 || c % heads) return -1;
  return (int64_t)batch * (int64_t)((size_t)heads * (t + 1) + 2 * (size_t)heads * c + c) * 4;
}

static int attnpool_check(const VqvsAttnPool* d, const char* what) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->c > 0 && d->t > 0 && d->heads > 0 && d->heads <= AP_THREADS && d->c % d->heads == 0 &&
                     d->c_out > 0, "%s: bad sizes", what);
  VQVS_CHECK_ARG(d->h && d->prep && d->w_qkv && d->b_qkv && d->w_proj && d->b_proj && d->ws, "%s: null pointer", what);
  return VQVS_OK;
}

extern "C" int vqvs_attnpool_fwd(const VqvsAttnPool* d, void* stream) {
  int rc = attnpool_check(d, "attnpool_fwd");
  if (rc) return rc;
  VQVS_CHECK_ARG(d->out, "attnpool_fwd: null output");
  attnpool_fwd_kernel<<<d->batch, AP_THREADS, 2 * d->c * sizeof(float), (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_attnpool_fwd");
  return VQVS_OK;
}

extern "C" int vqvs_attnpool_bwd(const VqvsAttnPool* d, void* stream) {
  int rc = attnpool_check(d, "attnpool_bwd");
  if (rc) return rc;
  VQVS_CHECK_ARG(d->d_out && d->d_act, "attnpool_bwd: null gradient pointer");
  const size_t smem = ((size_t)3 * d->c + (size_t)d->heads * d->c + (size_t)d->heads * (d->t + 1)) * sizeof(float);
  VQVS_CHECK_ARG(smem <= 48 * 1024, "attnpool_bwd: %zu bytes of shared memory needed", smem);
  attnpool_bwd_kernel<<<d->batch, AP_THREADS, smem, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_attnpool_bwd");
  return VQVS_OK;
}

extern "C" int vqvs_cls_head_fwd(const VqvsClsHead* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->dim > 0 && d->labels > 0 && d->stem && d->w && d->b && d->logits, "cls_head_fwd: bad arguments");
  cls_head_fwd_kernel<<<d->batch, 128, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_cls_head_fwd");
  return VQVS_OK;
}

extern "C" int vqvs_cls_head_bwd(const VqvsClsHead* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->dim > 0 && d->labels > 0 && d->stem && d->w && d->d_logits && d->d_stem, "cls_head_bwd: bad arguments");
  cls_head_bwd_kernel<<<d->batch, 256, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_cls_head_bwd");
  return VQVS_OK;
}
