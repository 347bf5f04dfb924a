/*
 * vqvs.h -- C ABI of libvqvs.so: the sm_100a kernels behind the diffusion-sampling
 * hot path of unixpickle/vq-voice-swap.
 *
 * The reference has no FFI: its boundary is the Python module API, and every op
 * below replaces a chain of ATen calls made by a reference function (cited per
 * entry as <file>:<lines>, paths relative to the reference's vq_voice_swap/).
 * The Python host (vq_voice_swap_b200/) binds these with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host_*; tensors are dense,
 *     row-major (NCT: batch, channel, time), fp32 unless stated
 *   - `stream` is a cudaStream_t passed as void*; the library never allocates,
 *     frees or synchronises; every call is asynchronous on `stream`
 *   - return 0 on success, negative VQVS_E* on failure; vqvs_last_error() gives
 *     a thread-local message for the last failure on the calling thread
 *   - per-channel statistics buffers are double[batch][channels][2] holding
 *     (sum, sum of squares) over time; producers ACCUMULATE into them with
 *     atomics, so the caller zeroes them once per forward (vqvs_run MEMSET op);
 *     a producer granted a statistics granularity G (VQVS_CONV_STAT_GRAN_SHIFT)
 *     adds the sums of G consecutive channels into the first channel's slot
 */
#ifndef VQVS_H_
#define VQVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQVS_ABI_VERSION 7

#define VQVS_OK 0
#define VQVS_EINVAL (-1)   /* bad argument / unsupported shape */
#define VQVS_ECUDA (-2)    /* CUDA runtime error (message has the string) */
#define VQVS_EARCH (-3)    /* device is not sm_100 and the op needs tcgen05 */

int vqvs_abi_version(void);
const char* vqvs_last_error(void);
/* cc_major*10+cc_minor and SM count of the current device. */
int vqvs_device_info(int* cc, int* sm_count);

/* ------------------------------------------------------------------------- */
/* Resize modes of reference models/unet.py:319-334 (Resize.forward).         */
#define VQVS_RESIZE_NONE 0
#define VQVS_RESIZE_DOWN2 1 /* F.avg_pool1d(x, 2) */
#define VQVS_RESIZE_UP2 2   /* nearest, out[i] = in[i/2] */

#define VQVS_SKIP_NONE 0
#define VQVS_SKIP_IDENTITY 1 /* out += resize(skip input)            unet.py:265-271,316 */
#define VQVS_SKIP_CONV1X1 2  /* out += W_skip * resize(skip input) + b_skip */

/*
 * One fused 1-D convolution = one half of a reference ResBlock
 * (models/unet.py:280-316), or cond_proj / encoder-out convs:
 *
 *   u[n,c,:]  = resize( act ? GELU(x[n,c,:]*scale[n,c] + shift[n,c]) : x[n,c,:] )
 *   out[n,o,t]= bias[o] + sum_{c,k} w[o,c,k] * u[n,c,t+(k-ksize/2)*dilation]   (zero padded)
 *             + skip term
 *   stats_out[n,o] += (sum_t out, sum_t out^2)          when stats_out != NULL
 *
 * x is the channel concatenation [xa ; xb] (torch.cat of unet.py:156 never
 * materialised); scale/shift carry GroupNorm (and FiLM, unet.py:311-314)
 * folded per (sample, channel) by vqvs_gn_finalize.  GELU is the erf form of
 * unet.py:341-342 (NOT tanh / SiLU), evaluated by approximations with <= 6.4e-7
 * absolute error (tools/fit_gelu_poly.py; common.cuh gelu_as).
 */
/* VqvsConv.reserved_ flag: the producer may accumulate statistics per channel PAIR (both channels' sums go to
 * the even channel's slot, the odd slot stays as the caller zeroed it).  Valid when every GroupNorm that
 * consumes them has an even number of channels per group (vqvs_gn_finalize sums over the group anyway). */
#define VQVS_CONV_PAIR_STATS 1024
/* bits 12..15 of VqvsConv.reserved_: log2(G), G = 2, 4, 8 or 16 -- the producer may merge the statistics of G
 * consecutive, G-aligned channels into the first channel's slot (requires VQVS_CONV_PAIR_STATS; 0 means G = 2).
 * Valid when every consuming GroupNorm group is a union of whole G-granules of this tensor. */
#define VQVS_CONV_STAT_GRAN_SHIFT 12
/* bits 16..17 of VqvsConv.reserved_: tensor-core operand format of vqvs_conv1d_umma; must match the format the weight
 * image was packed with (vqvs_pack_conv_weights).
 *   VQVS_PREC_BF16X3  activations and weights split into bf16 hi + lo, three products per tap (~2^-16 relative operand
 *                     error; unet64 forward 1.2e-5 vs the fp32 oracle) -- the default;
 *   VQVS_PREC_F16     one fp16 x fp16 product per tap (2^-11 operand rounding).  Meant for the deep, tensor-bound
 *                     levels only: with every conv of C_out >= 4*base_channels in this format a UNet forward stays at
 *                     6e-5 (tools/precision_study.py), a third of the tensor work of those layers. */
#define VQVS_CONV_PREC_SHIFT 16
#define VQVS_PREC_BF16X3 0
#define VQVS_PREC_F16 1

struct VqvsGnFinalize; /* defined below */

typedef struct VqvsConv {
  int32_t batch;
  int32_t c_a, c_b;        /* channels of xa and xb (c_b = 0: no concat) */
  int32_t t_in;            /* length of xa/xb */
  int32_t c_out, t_out;    /* t_out = t_in, t_in/2 or 2*t_in per `resize` */
  int32_t ksize;           /* 1 or 3 */
  int32_t dilation;
  int32_t resize;          /* VQVS_RESIZE_* applied to the activated input */
  int32_t act;             /* 1: scale/shift + GELU prologue; 0: raw input */
  int32_t skip_mode;       /* VQVS_SKIP_* */
  int32_t s_a, s_b;        /* channels of the skip sources sa, sb */
  int32_t t_skip;          /* length of sa/sb (the block input; differs from t_in in resize blocks) */
  int32_t skip_resize;     /* VQVS_RESIZE_* applied to the raw skip input: t_out = resize(t_skip) */
  int32_t reserved_;       /* flags: VQVS_CONV_PAIR_STATS; low bits are profiling switches (0 in production) */
  const float* xa;
  const float* xb;
  const float* scale;      /* [batch, c_a+c_b] */
  const float* shift;      /* [batch, c_a+c_b] */
  const float* w;          /* [c_out, c_a+c_b, ksize] fp32 (SIMT path) */
  const float* bias;       /* [c_out] */
  const float* sa;         /* skip sources, length t_skip each */
  const float* sb;
  const float* w_skip;     /* [c_out, s_a+s_b] fp32 (SIMT path) */
  const float* b_skip;     /* [c_out] */
  const void* w_packed;    /* tcgen05 operand image made by vqvs_pack_conv_weights (UMMA path) */
  float* out;              /* [batch, c_out, t_out] */
  double* stats_out;       /* [batch, c_out, 2] or NULL */
  /* HOST pointer or NULL.  When set (UMMA path, act = 1) the kernel derives scale/shift for its input itself from
   * the producers' statistics, exactly as vqvs_gn_finalize would (same fp64 formulas), so the separate finalize
   * launch between two convs disappears; `scale` / `shift` are then ignored. */
  const struct VqvsGnFinalize* gn;
} VqvsConv;

/* fp32 CUDA-core implementation (any shape). */
int vqvs_conv1d_fused(const VqvsConv* d, void* stream);
/* tcgen05 / TMEM implementation (bf16x3 split operands, fp32 accumulate);
 * needs channel counts that are multiples of 16 and c_out <= 512. */
int vqvs_conv1d_umma(const VqvsConv* d, void* stream);
/* 1 if vqvs_conv1d_umma accepts this descriptor. */
int vqvs_conv1d_umma_supported(const VqvsConv* d);
/* Bytes of the packed operand image for a conv (main taps + optional 1x1 skip) in operand format `prec`
 * (VQVS_PREC_BF16X3 / VQVS_PREC_F16); -1 if the shape is not supported. */
int64_t vqvs_packed_weight_bytes(int c_out, int c_in, int ksize, int c_skip, int prec);
/* Build the image from fp32 weights w[c_out,c_in,ksize], w_skip[c_out,c_skip] (device). */
int vqvs_pack_conv_weights(const float* w, const float* w_skip, int c_out, int c_in, int ksize,
                           int c_skip, int prec, void* packed, void* stream);

/*
 * GroupNorm statistics -> per-(sample, channel) affine, with optional FiLM:
 *   mean_g, var_g over the (C/G)*T elements of group g (biased var, eps = 1e-5)
 *   s = rstd_g*gamma[c];  h = beta[c] - mean_g*s
 *   FiLM (unet.py:311-314):  s' = s*(1+a[n,c]); h' = h*(1+a[n,c]) + b[n,c]
 * replaces nn.GroupNorm of models/unet.py:345-349 (stats come from producers).
 * The normalised tensor is the concat of two sources with c_a and c_b channels.
 */
typedef struct VqvsGnFinalize {
  int32_t batch, c_a, c_b, groups;
  int64_t count;               /* elements per channel (T) */
  const double* stats_a;       /* [batch, c_a, 2] */
  const double* stats_b;       /* [batch, c_b, 2] or NULL */
  const float* gamma;          /* [c_a+c_b] */
  const float* beta;
  const float* film;           /* NULL or &ab[0][offset]: a at [n*film_stride + c], b at [.. + C + c] */
  int64_t film_stride;
  float* scale;                /* [batch, c_a+c_b] */
  float* shift;
} VqvsGnFinalize;
int vqvs_gn_finalize(const VqvsGnFinalize* d, void* stream);

/* Standalone per-channel (sum, sumsq) of x[batch, c, t] accumulated into stats
 * (used for tensors not produced by our convs, e.g. user-supplied inputs). */
int vqvs_channel_stats(const float* x, int batch, int c, int t, double* stats, void* stream);

/*
 * in_conv (+ conditioning):   models/unet.py:137-139, :229
 *   out[n,o,t] = b[o] + sum_k w[o,0,k]*x[n,0,t+k-1]  (+ cond[n,o,floor(t*t_cond/t)])
 * and accumulates stats_out.
 */
typedef struct VqvsConvIn {
  int32_t batch, c_out, t, t_cond;
  const float* x;      /* [batch, 1, t] */
  const float* w;      /* [c_out, 1, 3] */
  const float* bias;   /* [c_out] */
  const float* cond;   /* NULL or [batch, c_out, t_cond] = cond_proj(cond) */
  float* out;
  double* stats_out;
} VqvsConvIn;
int vqvs_conv_in(const VqvsConvIn* d, void* stream);

/*
 * Final GN -> GELU -> Conv1d(c -> 1, k3) (models/unet.py:112-116,162) with the
 * DDPM update of diffusion/diffusion.py:48-90 fused into the store:
 *   mode EPS      : out = eps
 *   mode PREV     : out = c1[n]*(x_t - c2[n]*eps) + sigma[n]*noise        (:69-70, :90)
 *   mode X0_SUM   : out = eps and x0_sum[n] += sum_t (x_t - c3[n]*eps)*c4[n]  (:85-87 first half)
 * coef is float[batch][8] = {c1=alpha^-1/2, c2=beta*(1-abar_t)^-1/2, sigma, c3=(1-abar_t)^1/2,
 * c4=abar_t^-1/2, c5=abar_t^1/2, c6=(1-abar_t)^-1/2, unused}; the host computes it in fp32 with the
 * reference's own operation order.
 */
#define VQVS_OUT_EPS 0
#define VQVS_OUT_PREV 1
#define VQVS_OUT_X0_SUM 2
typedef struct VqvsConvOut {
  int32_t batch, c_in, t, mode;
  const float* x;       /* [batch, c_in, t] */
  const float* scale;   /* [batch, c_in] (from vqvs_gn_finalize) */
  const float* shift;
  const float* w;       /* [1, c_in, 3] */
  const float* bias;    /* [1] */
  const float* x_t;     /* [batch, 1, t]  (modes PREV, X0_SUM) */
  const float* noise;   /* [batch, 1, t] or NULL = zeros (mode PREV) */
  const float* coef;    /* [batch, 8] */
  float* out;           /* [batch, 1, t] */
  double* x0_sum;       /* [batch] (mode X0_SUM) */
} VqvsConvOut;
int vqvs_conv_out(const VqvsConvOut* d, void* stream);

/*
 * Elementwise DDPM finisher for constrain / cond_fn (diffusion/diffusion.py:80-90):
 *   if use_x0_mean: x0 = clamp((x_t - c3*eps)*c4 - x0_sum[n]/t, -1, 1); eps = (x_t - x0*c5)*c6
 *   out = c1*(x_t - c2*eps) + sigma*noise
 */
typedef struct VqvsDdpmFinish {
  int32_t batch, t, use_x0_mean;
  const float* x_t;
  const float* eps;
  const float* noise;    /* NULL = zeros */
  const float* coef;     /* [batch, 8] as above */
  const double* x0_sum;  /* [batch] */
  float* out;
} VqvsDdpmFinish;
int vqvs_ddpm_finish(const VqvsDdpmFinish* d, void* stream);
/* x0_sum[n] += sum_t (x_t - c3*eps)*c4   (for an eps that did not come from vqvs_conv_out) */
int vqvs_ddpm_x0_sum(const float* x_t, const float* eps, const float* coef, int batch, int t,
                     double* x0_sum, void* stream);

/* ------------------------------------------------------------------------- */
/* Classifier guidance (BASELINE configs[4]): reference models/classifier.py:18-191 evaluated forward AND backward
 * (input gradient only) without ATen/autograd -- sample_diffusion.py:34-42 needs d log p(label | x_t, t) / d x_t.
 * The stem's convolutions, including the transposed ("dgrad") ones, are vqvs_conv1d_umma launches; these ops are the
 * pieces between them.  A GroupNorm(+FiLM) followed by GELU, v = gelu(z*S + H), is differentiated in three steps that
 * mirror the forward's producer-statistics / consumer-finalize split:
 *   vqvs_gn_bwd_prep      prep = [S | H | mean | rstd | gf] per (n, c) from the forward's statistics (gf = gamma*(1+a))
 *   vqvs_gelu_bwd         q = d_in * gelu'(z*S + H) * gf;  acc[n,c] += (sum_t q, sum_t q*zhat),  zhat = (z-mean)*rstd
 *   vqvs_gn_bwd_finalize  coef = [A | B | C] with dz = A*q + B*z + C = rstd*(q - mean_g(q) - zhat*mean_g(q*zhat))
 *   vqvs_affine3          out = A*q + B*z + C (+ add), add = the skip path's gradient (optionally through avg_pool^T)
 */
typedef struct VqvsGnBwdPrep { const VqvsGnFinalize* gn; float* prep; /* [5][batch*C] */ } VqvsGnBwdPrep;
int vqvs_gn_bwd_prep(const VqvsGnFinalize* d, float* prep, void* stream);
typedef struct VqvsGeluBwd {
  int32_t batch, c, t;
  int32_t up;                /* resize of the FORWARD between gelu(GN(z)) and the conv: 0 none; 1 avg_pool1d (d_in is t/2 long,
                                reaches i as 0.5*d_in[i/2]); 2 nearest x2 (d_in is 2t long, i collects d_in[2i] + d_in[2i+1]) */
  int32_t c_total, c_off;    /* the GroupNorm spans c_total concatenated channels, this source is [c_off, c_off + c): prep,
                                acc and d_in [batch, c_total, .] are indexed with them; z and q are dense [batch, c, t] */
  const float* d_in; const float* z; const float* prep;
  float* q;                  /* [batch,c,t] */
  double* acc;               /* [batch,c,2], zeroed by the caller */
} VqvsGeluBwd;
int vqvs_gelu_bwd(const VqvsGeluBwd* d, void* stream);
typedef struct VqvsGnBwdFinalize {
  int32_t batch, c, groups, pad_;
  int64_t count;             /* positions per channel */
  const double* acc; const float* prep;
  float* coef;               /* [3][batch*c] */
} VqvsGnBwdFinalize;
int vqvs_gn_bwd_finalize(const VqvsGnBwdFinalize* d, void* stream);
typedef struct VqvsAffine3 {
  int32_t batch, c, t;
  int32_t add_mode;          /* 0: no add, 1: + add[i], 2: + 0.5*add[i/2] (avg_pool^T), 3: + add[2i] + add[2i+1] (nearest x2 ^T) */
  int32_t c_total, c_off;    /* coef and add [batch, c_total, .] are indexed like VqvsGeluBwd's d_in */
  const float* q; const float* z; const float* coef; const float* add;
  const float* add2;         /* optional dense [batch, c, t]: a gradient this tensor already received from another consumer */
  float* out;
} VqvsAffine3;
int vqvs_affine3(const VqvsAffine3* d, void* stream);
/* backward = 0: dst[row, j] = src[row, j*rate] (F.interpolate(h, size = t/rate), nearest -- models/encoder_predictor.py:55-57);
 * backward = 1: dst[row, j*rate] = src[row, j], zero elsewhere.  t is the LONG length. */
int vqvs_stride_sample(const float* src, float* dst, int rows, int t, int rate, int backward, void* stream);
/* Input gradient of Conv1d(1 -> c, k = 3, pad 1) (models/classifier.py:79): dx[n,t] = sum_{c,k} w[c,0,k]*dh[n,c,t+1-k]. */
typedef struct VqvsConvInBwd { int32_t batch, c, t, pad_; const float* dh; const float* w; float* dx; } VqvsConvInBwd;
int vqvs_conv_in_bwd(const VqvsConvInBwd* d, void* stream);
/*
 * AttentionPool1d (models/classifier.py:133-191) on tokens [0 ; gelu(GN(h))]: 1x1 qkv projection, per-head softmax
 * attention of QUERY ROW 0 only (the only row the reference keeps, :158), 1x1 output projection.  Token 0 is the zero pad,
 * so the query is the projection's bias.  `prep` is the vqvs_gn_bwd_prep block of the GroupNorm in front (S, H are used).
 * ws: vqvs_attnpool_workspace_bytes(batch, c, t, heads) bytes, written by fwd and read by bwd.
 * bwd writes d_act[batch,c,t] = gradient w.r.t. gelu(GN(h)), to be fed to vqvs_gelu_bwd as d_in.
 */
typedef struct VqvsAttnPool {
  int32_t batch, c, t, heads, c_out, pad_;
  const float* h; const float* prep;
  const float* w_qkv; const float* b_qkv;    /* [3c, c], [3c] */
  const float* w_proj; const float* b_proj;  /* [c_out, c], [c_out] */
  float* ws;
  float* out;             /* fwd: [batch, c_out] */
  const float* d_out;     /* bwd: [batch, c_out] */
  float* d_act;           /* bwd: [batch, c, t] */
} VqvsAttnPool;
int64_t vqvs_attnpool_workspace_bytes(int batch, int c, int t, int heads);
int vqvs_attnpool_fwd(const VqvsAttnPool* d, void* stream);
int vqvs_attnpool_bwd(const VqvsAttnPool* d, void* stream);
/* Classifier head (models/classifier.py:28-45): logits = W*gelu(stem) + b;  d_stem = gelu'(stem) * W^T d_logits. */
typedef struct VqvsClsHead {
  int32_t batch, dim, labels, pad_;
  const float* stem; const float* w; const float* b;
  float* logits; const float* d_logits; float* d_stem;
} VqvsClsHead;
int vqvs_cls_head_fwd(const VqvsClsHead* d, void* stream);
int vqvs_cls_head_bwd(const VqvsClsHead* d, void* stream);

/* ------------------------------------------------------------------------- */
/* ConvMFCCEncoder front end (reference models/conv_encoder.py:14-147, version 1 -- the encoder of the published
 * vqvae-unet-mfcc checkpoint): inverse mu-law (:146-147) -> torchaudio MFCC (n_fft = 2*hop Hann frames, centre/reflect,
 * power spectrum, mel filter bank, log(mel + 1e-6), DCT) -> deltas and delta-deltas (:136-143).
 * window [n_fft], cos_t / sin_t [n_bins, n_fft] (DFT basis), fb [n_bins, n_mels], dct [n_mels, n_mfcc]: device copies of
 * the module's buffers.  mfcc: scratch [batch, n_mfcc, frames]; out: [batch, c_pad, frames], channels >= 3*n_mfcc zeroed
 * (c_pad rounds 39 up to the conv kernel's 16-channel granularity). */
typedef struct VqvsMfcc {
  int32_t batch, t, n_fft, hop, n_bins, n_mels, n_mfcc, frames, c_pad, ulaw;
  const float* x; const float* window; const float* cos_t; const float* sin_t; const float* fb; const float* dct;
  float* mfcc; float* out;
} VqvsMfcc;
int vqvs_mfcc39(const VqvsMfcc* d, void* stream);
/* out[row, i] = (res ? res[row, i] : 0) + GELU(h[row, i]), i < t: the "conv -> GELU (-> + x)" blocks of
 * conv_encoder.py:63-86,121-133 (rows = batch*channels; pitches in floats). */
int vqvs_gelu_add(const float* h, int h_pitch, const float* res, int res_pitch, float* out, int out_pitch, int rows, int t,
                  void* stream);
/* even[row, j] = x[row, 2j], odd[row, j] = x[row, 2j+1] (zero past the end): turns the stride-2, k = 4 conv of
 * conv_encoder.py:68-73 into a k = 3 conv over the channel concatenation [even ; odd]. */
int vqvs_deinterleave2(const float* x, int rows, int t, float* even, float* odd, int t_half, void* stream);

/*
 * Standard-normal noise keyed by (seed, GLOBAL sample index, step) for batch-sharded sampling: out[r, :]
 * (rows of `length` floats) depends only on (seed, first_row + r, step) -- Philox4x32-10 + Box-Muller on the device.
 * It replaces the reference's `torch.randn_like(x_t)` (diffusion/diffusion.py:62-63) where the result must not depend
 * on how a batch is split over GPUs; step = -1 is used for x_T.
 */
int vqvs_keyed_normal(float* out, int rows, int64_t length, uint64_t seed, int64_t first_row, int32_t step, void* stream);

/*
 * Timestep embedding (models/wavegrad.py:359-373, models/unet.py:40-45,133-135):
 *   e = [cos(t*f) | sin(t*f)];  emb = W2*GELU(W1*e + b1) + b2 (+ class_embed[label])
 * writes emb and GELU(emb) (the input of every FiLM Linear, unet.py:274-278).
 */
typedef struct VqvsTimeEmbed {
  int32_t batch, dim;        /* dim = 4*base_channels */
  const float* ts;           /* [batch] */
  const float* freqs;        /* [dim/2], built by the host exactly like the reference */
  const float* w1; const float* b1; const float* w2; const float* b2; /* [dim,dim], [dim] */
  const float* class_embed;  /* NULL or [num_labels, dim] */
  const int64_t* labels;     /* NULL or [batch] */
  float* emb;                /* [batch, dim] */
  float* gelu_emb;           /* [batch, dim] */
} VqvsTimeEmbed;
int vqvs_time_embed(const VqvsTimeEmbed* d, void* stream);

/* out[i] = GELU(in[i]) (erf form, approximated to <= 6.4e-7 absolute); input of cond_layers when a caller supplies emb directly. */
int vqvs_gelu(const float* in, float* out, int64_t n, void* stream);

/* All FiLM Linear layers of a network in one launch: ab[n, :] = W_cat * gelu_emb[n] + b_cat,
 * W_cat = rows of every block's cond_layers.1.weight stacked (unet.py:277). */
int vqvs_film_linear(const float* gelu_emb, const float* w_cat, const float* b_cat, int batch,
                     int dim, int n_out, float* ab, void* stream);

/*
 * VQ nearest-codebook search (vq.py:127-131, 199-221): x[n, c, t1], dict[d, c] ->
 * idx int64 [n, t1] = argmin_d fl(fl(-2*dot + |dict_d|^2) + |x|^2), first minimum wins.
 * Dots and norms are accumulated in fp64 and rounded once to fp32.
 */
int vqvs_vq_argmin(const float* x, const float* dict, int n, int c, int t1, int d, int64_t* idx,
                   void* stream);
/* vq.py:98-110: out[n, c, t1] = dict[idx[n, t1], c] */
int vqvs_vq_embed(const int64_t* idx, const float* dict, int n, int c, int t1, int d, float* out,
                  void* stream);

/* ------------------------------------------------------------------------- */
/* Programs: a forward pass is a static list of ops; vqvs_run launches them    */
/* in order on one stream with one host call (replaces ~900 ATen dispatches    */
/* per reference UNetPredictor.forward, models/unet.py:118-163).               */
#define VQVS_OP_CONV_SIMT 1
#define VQVS_OP_CONV_UMMA 2
#define VQVS_OP_GN_FINALIZE 3
#define VQVS_OP_CONV_IN 4
#define VQVS_OP_CONV_OUT 5
#define VQVS_OP_TIME_EMBED 6
#define VQVS_OP_FILM 7
#define VQVS_OP_MEMSET 8
#define VQVS_OP_DDPM_FINISH 9
#define VQVS_OP_GN_BWD_PREP 10
#define VQVS_OP_GELU_BWD 11
#define VQVS_OP_GN_BWD_FINALIZE 12
#define VQVS_OP_AFFINE3 13
#define VQVS_OP_CONV_IN_BWD 14
#define VQVS_OP_ATTNPOOL_FWD 15
#define VQVS_OP_ATTNPOOL_BWD 16
#define VQVS_OP_CLS_HEAD_FWD 17
#define VQVS_OP_CLS_HEAD_BWD 18

typedef struct VqvsFilm {
  const float* gelu_emb; const float* w_cat; const float* b_cat;
  int32_t batch, dim, n_out; float* ab;
} VqvsFilm;
typedef struct VqvsMemset { void* ptr; int64_t bytes; } VqvsMemset;

typedef struct VqvsOp {
  int32_t kind;
  const void* desc; /* host pointer to the matching Vqvs* struct */
} VqvsOp;
/* Returns 0, or the first failing op's status (message names the op index). */
int vqvs_run(const VqvsOp* ops, int n_ops, void* stream);
/* Same, bracketing every op with CUDA events on `stream`; host_ms[i] receives op i's device time
 * in milliseconds (synchronises the stream at the end; measurement aid for bench.py). */
int vqvs_run_timed(const VqvsOp* ops, int n_ops, void* stream, float* host_ms);

/* Device workspace (bytes) op `kind` (VQVS_OP_*) needs BEYOND the buffers its descriptor names (the library
 * never allocates, callers size workspaces with this).  Every op works out of its descriptor's buffers and shared
 * memory -- 0 -- except the attention pool, whose VqvsAttnPool.ws holds vqvs_attnpool_workspace_bytes(batch, c, t, heads)
 * bytes shared by its forward and backward; the packed conv weight image is a separate, persistent buffer
 * (vqvs_packed_weight_bytes).  Returns -1 for an unknown kind or a null descriptor. */
int64_t vqvs_workspace_bytes(int kind, const void* desc);

/* Ops executed by vqvs_run / vqvs_run_timed since the library was loaded, indexed by VQVS_OP_* (32 slots; out32[VQVS_OP_CONV_UMMA] =
 * tcgen05 conv launches, ...; memsets are counted under VQVS_OP_MEMSET).  Evidence hook for the script-level tests and
 * bench.py's gpu_launches: a run that went through a fallback would leave these at zero. */
int vqvs_launch_counts(unsigned long long* out32);

/* Role profiler of the last vqvs_conv1d_umma launched with debug flag 512 (cycles per phase of CTA 0):
 * [0..3] transform warp 0: wait operand slot, wait raw, work, loop; [4..7] TMA; [8..11] MMA; [12..15] epilogue. */
int vqvs_debug_prof(unsigned long long* host32);
/* Shared-memory / pipeline plan vqvs_conv1d_umma would use for `d` (tuning aid and test hook, no launch):
 * out16 = {n_tiles, n_tile, stack, w_resident, kbs, mt, nbuf, ab_slots, ab_slot_bytes, raw_slots, raw_slot_bytes,
 *          smem_bytes, tma, main_stages, skip_stages, b_slots}. */
int vqvs_debug_geo(const VqvsConv* d, int* out16);

/* tcgen05 self-test: runs D[128,n] = A[128,k] * B[n,k]^T through the exact smem layout,
 * descriptors and TMEM read-back used by vqvs_conv1d_umma, with A rows shifted by `row_shift`.
 * a: [128+row_shift, k] fp32, b: [n, k] fp32, d: [128, n] fp32 (device).
 * variant 0 = production (bf16x3); 1 = LBO/SBO swapped (diagnostic); 2 = hi*hi only (plain bf16). */
int vqvs_umma_selftest(const float* a, const float* b, float* d, int n, int k, int row_shift,
                       int variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQVS_H_ */
