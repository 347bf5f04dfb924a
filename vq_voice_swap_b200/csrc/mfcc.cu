// Front end of reference models/conv_encoder.py (ConvMFCCEncoder, the encoder of the published vqvae-unet-mfcc
// checkpoint): inverse mu-law -> MFCC (torchaudio.transforms.MFCC, version 1: STFT n_fft = 2*hop, Hann window, centre +
// reflect padding, power spectrum, HTK mel filter bank, log(mel + 1e-6), orthonormal DCT-II) -> first and second
// differences (conv_encoder.py:136-143) -> 39 channels, plus the two small elementwise helpers its conv stack needs
// (GELU + residual after a conv; even/odd split for the stride-2 conv).  Once per encode: latency-, not throughput-bound.
#include "common.cuh"

namespace vqvs {

// One CTA per (frame, sample).  window / cos / sin tables, filter bank and DCT matrix come from the module's own buffers.
__global__ void __launch_bounds__(256) mfcc_kernel(VqvsMfcc d) {
  extern __shared__ float sm[];  // frame[n_fft] | power[n_bins] | mel[n_mels]
  float* frame = sm;
  float* power = frame + d.n_fft;
  float* mel = power + d.n_bins;
  const int f = blockIdx.x, n = blockIdx.y;
  const float* x = d.x + (size_t)n * d.t;
  const int start = f * d.hop - d.n_fft / 2;  // center = True
  for (int i = threadIdx.x; i < d.n_fft; i += blockDim.x) {
    int p = start + i;
    if (p < 0) p = -p;                          // reflect padding (no edge repeat)
    if (p >= d.t) p = 2 * (d.t - 1) - p;
    float v = x[p];
    if (d.ulaw) {  // invert_ulaw, conv_encoder.py:146-147
      const float a = fabsf(v);
      v = copysignf((powf(256.0f, a) - 1.0f) * (1.0f / 255.0f), v);
    }
    frame[i] = v * d.window[i];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d.n_bins; k += blockDim.x) {
    const float* c = d.cos_t + (size_t)k * d.n_fft;
    const float* s = d.sin_t + (size_t)k * d.n_fft;
    float re = 0.f, im = 0.f;
    for (int i = 0; i < d.n_fft; ++i) {
      re = fmaf(frame[i], c[i], re);
      im = fmaf(frame[i], s[i], im);
    }
    power[k] = re * re + im * im;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < d.n_mels; m += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < d.n_bins; ++k) a = fmaf(power[k], d.fb[(size_t)k * d.n_mels + m], a);
    mel[m] = logf(a + 1e-6f);  // log_mels = True
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d.n_mfcc; c += blockDim.x) {
    float a = 0.f;
    for (int m = 0; m < d.n_mels; ++m) a = fmaf(mel[m], d.dct[(size_t)m * d.n_mfcc + c], a);
    d.mfcc[((size_t)n * d.n_mfcc + c) * d.frames + f] = a;
  }
}

// out[n, 0:13] = mfcc, [13:26] = deltas(mfcc), [26:39] = deltas(deltas), channels [39, c_pad) = 0
__device__ __forceinline__ float delta_at(const float* r, int f, int frames) {  // ((r[f-1] - r[f]) + (r[f] - r[f+1])) / 2, edges repeat
  const float l = r[f > 0 ? f - 1 : 0], c = r[f], rr = r[f + 1 < frames ? f + 1 : frames - 1];
  return ((l - c) + (c - rr)) * 0.5f;
}
__global__ void mfcc_deltas_kernel(VqvsMfcc d) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, n = blockIdx.z;
  if (f >= d.frames) return;
  float* out = d.out + (size_t)n * d.c_pad * d.frames;
  if (c >= 3 * d.n_mfcc) {
    out[(size_t)c * d.frames + f] = 0.f;
    return;
  }
  const int base = c % d.n_mfcc, order = c / d.n_mfcc;
  const float* r = d.mfcc + ((size_t)n * d.n_mfcc + base) * d.frames;
  float v;
  if (order == 0) v = r[f];
  else if (order == 1) v = delta_at(r, f, d.frames);
  else {  // deltas of the delta sequence, whose own edges follow the same rule
    auto dl = [&](int g) { return delta_at(r, g, d.frames); };
    const float l = dl(f > 0 ? f - 1 : 0), cc = dl(f), rr = dl(f + 1 < d.frames ? f + 1 : d.frames - 1);
    v = ((l - cc) + (cc - rr)) * 0.5f;
  }
  out[(size_t)c * d.frames + f] = v;
}

// out[row, i] = (res ? res[row, i] : 0) + gelu(h[row, i]) for i < t, with independent row pitches
__global__ void gelu_add_kernel(const float* __restrict__ h, int h_pitch, const float* res, int res_pitch, float* out,
                                int out_pitch, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = blockIdx.y;
  if (i >= t) return;
  float v = gelu_erf(h[row * h_pitch + i]);
  if (res) v += res[row * res_pitch + i];
  out[row * out_pitch + i] = v;
}

// even[row, j] = x[row, 2j], odd[row, j] = x[row, 2j+1] (zero past the end), j < t_half
__global__ void deinterleave2_kernel(const float* __restrict__ x, int t, float* __restrict__ even, float* __restrict__ odd,
                                     int t_half) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t row = blockIdx.y;
  if (j >= t_half) return;
  even[row * t_half + j] = 2 * j < t ? x[row * t + 2 * j] : 0.f;
  odd[row * t_half + j] = 2 * j + 1 < t ? x[row * t + 2 * j + 1] : 0.f;
}

}  // namespace vqvs

using namespace vqvs;

extern "C" int vqvs_mfcc39(const VqvsMfcc* d, void* stream) {
  VQVS_CHECK_ARG(d && d->batch > 0 && d->t > 0 && d->n_fft > 0 && d->hop > 0 && d->n_bins == d->n_fft / 2 + 1 && d->n_mels > 0 &&
                     d->n_mfcc > 0 && d->frames == d->t / d->hop + 1 && d->c_pad >= 3 * d->n_mfcc, "mfcc39: bad geometry");
  VQVS_CHECK_ARG(d->t > d->n_fft / 2, "mfcc39: reflect padding needs more than n_fft/2 samples");
  VQVS_CHECK_ARG(d->x && d->window && d->cos_t && d->sin_t && d->fb && d->dct && d->mfcc && d->out, "mfcc39: null pointer");
  const size_t smem = (size_t)(d->n_fft + d->n_bins + d->n_mels) * sizeof(float);
  VQVS_CHECK_ARG(smem <= 48 * 1024, "mfcc39: n_fft %d too large", d->n_fft);
  mfcc_kernel<<<dim3(d->frames, d->batch), 256, smem, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_mfcc39 (spectrum)");
  mfcc_deltas_kernel<<<dim3(ceil_div(d->frames, 128), d->c_pad, d->batch), 128, 0, (cudaStream_t)stream>>>(*d);
  VQVS_CHECK_LAUNCH("vqvs_mfcc39 (deltas)");
  return VQVS_OK;
}

extern "C" int vqvs_gelu_add(const float* h, int h_pitch, const float* res, int res_pitch, float* out, int out_pitch, int rows,
                             int t, void* stream) {
  VQVS_CHECK_ARG(h && out && rows > 0 && rows <= 65535 && t > 0 && h_pitch >= t && out_pitch >= t && (!res || res_pitch >= t),
                 "gelu_add: bad arguments");
  gelu_add_kernel<<<dim3(ceil_div(t, 128), rows), 128, 0, (cudaStream_t)stream>>>(h, h_pitch, res, res_pitch, out, out_pitch, t);
  VQVS_CHECK_LAUNCH("vqvs_gelu_add");
  return VQVS_OK;
}

extern "C" int vqvs_deinterleave2(const float* x, int rows, int t, float* even, float* odd, int t_half, void* stream) {
  VQVS_CHECK_ARG(x && even && odd && rows > 0 && rows <= 65535 && t > 0 && t_half >= (t + 1) / 2, "deinterleave2: bad arguments");
  deinterleave2_kernel<<<dim3(ceil_div(t_half, 128), rows), 128, 0, (cudaStream_t)stream>>>(x, t, even, odd, t_half);
  VQVS_CHECK_LAUNCH("vqvs_deinterleave2");
  return VQVS_OK;
}
