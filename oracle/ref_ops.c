/*
 * TEST INFRASTRUCTURE -- plain C restatement of the ATen primitives the reference's hot path
 * dispatches to (the arithmetic lives in PyTorch, a third-party dependency that is absent from
 * /root/reference and unpinned in its setup.py:6; torch 2.11.0 in the build container).
 * Call sites restated: F.conv1d (reference models/unet.py:47,49,115,267,283,287), nn.GroupNorm
 * (unet.py:345-349), nn.GELU exact (unet.py:341-342), F.avg_pool1d / nearest interpolate
 * (unet.py:324-334), the VQ distance + argmin expression (vq.py:212-221,131) and the DDPM update
 * (diffusion/diffusion.py:48-90).  Single precision, sequential summation; checked against
 * oracle/hotpath.py (torch CPU) on small cases by tests/test_oracle_c.py.  Never linked by the product.
 */
#include <math.h>
#include <stdint.h>

/* y[n,o,t] = b[o] + sum_{c,k} w[o,c,k] * x[n,c,t + (k - K/2)*dil], zero padding */
void ref_conv1d(const float* x, const float* w, const float* b, float* y, int n, int c_in, int c_out, int t, int k, int dil) {
  for (int i = 0; i < n; ++i)
    for (int o = 0; o < c_out; ++o)
      for (int p = 0; p < t; ++p) {
        float acc = b ? b[o] : 0.f;
        for (int c = 0; c < c_in; ++c)
          for (int j = 0; j < k; ++j) {
            const int q = p + (j - k / 2) * dil;
            if (q >= 0 && q < t) acc += w[(o * c_in + c) * k + j] * x[((long)i * c_in + c) * t + q];
          }
        y[((long)i * c_out + o) * t + p] = acc;
      }
}

/* GroupNorm: per (sample, group) mean / biased variance over (C/G)*T, eps 1e-5, affine */
void ref_group_norm(const float* x, const float* gamma, const float* beta, float* y, int n, int c, int t, int groups) {
  const int cg = c / groups;
  for (int i = 0; i < n; ++i)
    for (int g = 0; g < groups; ++g) {
      double s = 0, ss = 0;
      for (int cc = g * cg; cc < (g + 1) * cg; ++cc)
        for (int p = 0; p < t; ++p) {
          const double v = x[((long)i * c + cc) * t + p];
          s += v;
          ss += v * v;
        }
      const double cnt = (double)cg * t, mean = s / cnt;
      const double rstd = 1.0 / sqrt(ss / cnt - mean * mean + 1e-5);
      for (int cc = g * cg; cc < (g + 1) * cg; ++cc)
        for (int p = 0; p < t; ++p) {
          const long idx = ((long)i * c + cc) * t + p;
          y[idx] = (float)((x[idx] - mean) * rstd) * gamma[cc] + beta[cc];
        }
    }
}

void ref_gelu(const float* x, float* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = 0.5f * x[i] * (1.0f + erff(x[i] * 0.70710678118654752440f));
}

void ref_avg_pool2(const float* x, float* y, int rows, int t) {
  for (int r = 0; r < rows; ++r)
    for (int p = 0; p < t / 2; ++p) y[(long)r * (t / 2) + p] = 0.5f * (x[(long)r * t + 2 * p] + x[(long)r * t + 2 * p + 1]);
}

void ref_upsample_nearest(const float* x, float* y, int rows, int t_in, int t_out) {
  const float scale = (float)t_in / (float)t_out;
  for (int r = 0; r < rows; ++r)
    for (int p = 0; p < t_out; ++p) {
      int s = (int)floorf((float)p * scale);
      if (s > t_in - 1) s = t_in - 1;
      y[(long)r * t_out + p] = x[(long)r * t_in + s];
    }
}

/* idx[v] = argmin_d ((-2*dot + |dict_d|^2) + |x_v|^2), first minimum; x is [n, c, t1] */
void ref_vq_argmin(const float* x, const float* dict, int64_t* idx, int n, int c, int t1, int d) {
  for (int i = 0; i < n; ++i)
    for (int p = 0; p < t1; ++p) {
      float xn = 0.f;
      for (int ch = 0; ch < c; ++ch) xn += x[((long)i * c + ch) * t1 + p] * x[((long)i * c + ch) * t1 + p];
      float best = INFINITY;
      int64_t arg = 0;
      for (int e = 0; e < d; ++e) {
        float dot = 0.f, dn = 0.f;
        for (int ch = 0; ch < c; ++ch) {
          dot += dict[(long)e * c + ch] * x[((long)i * c + ch) * t1 + p];
          dn += dict[(long)e * c + ch] * dict[(long)e * c + ch];
        }
        const float dist = (-2.f * dot + dn) + xn;
        if (dist < best) {
          best = dist;
          arg = e;
        }
      }
      idx[(long)i * t1 + p] = arg;
    }
}

/* One reverse-diffusion step for one sample (constrain=0, cond_fn=None): scalars follow diffusion.py:64-78 */
void ref_ddpm_step(const float* x_t, const float* eps, const float* noise, float* out, long len, float abar_t, float abar_prev,
                   int sigma_large) {
  const float alpha = abar_t / abar_prev, beta = 1.f - alpha;
  const float c1 = 1.f / sqrtf(alpha), c2 = beta * (1.f / sqrtf(1.f - abar_t));
  const float sig2 = sigma_large ? beta : beta * (1.f - abar_prev) / (1.f - abar_t);
  const float sigma = sqrtf(sig2);
  for (long i = 0; i < len; ++i) out[i] = c1 * (x_t[i] - c2 * eps[i]) + sigma * noise[i];
}
