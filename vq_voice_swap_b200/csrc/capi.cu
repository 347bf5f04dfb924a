// libvqvs: error channel, device query and the program runner.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vqvs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int run_film(const VqvsFilm* f, void* stream);

// kernels launched through vqvs_run / vqvs_run_timed since the library was loaded, per op kind (vqvs_launch_counts)
static unsigned long long g_launches[32];

}  // namespace vqvs

extern "C" int vqvs_abi_version(void) { return VQVS_ABI_VERSION; }

extern "C" const char* vqvs_last_error(void) { return vqvs::g_err; }

extern "C" int vqvs_device_info(int* cc, int* sm_count) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    vqvs::set_error("vqvs_device_info: %s", cudaGetErrorString(e));
    return VQVS_ECUDA;
  }
  if (cc) *cc = p.major * 10 + p.minor;
  if (sm_count) *sm_count = p.multiProcessorCount;
  return VQVS_OK;
}

static int run_impl(const VqvsOp* ops, int n_ops, void* stream, cudaEvent_t* ev);

extern "C" int vqvs_run(const VqvsOp* ops, int n_ops, void* stream) { return run_impl(ops, n_ops, stream, nullptr); }

extern "C" int vqvs_run_timed(const VqvsOp* ops, int n_ops, void* stream, float* host_ms) {
  if (!ops || n_ops <= 0 || !host_ms) {
    vqvs::set_error("vqvs_run_timed: bad arguments");
    return VQVS_EINVAL;
  }
  cudaEvent_t* ev = new cudaEvent_t[n_ops + 1];
  for (int i = 0; i <= n_ops; ++i) cudaEventCreate(&ev[i]);
  int rc = run_impl(ops, n_ops, stream, ev);
  if (rc == VQVS_OK) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) {
      vqvs::set_error("vqvs_run_timed: %s", cudaGetErrorString(e));
      rc = VQVS_ECUDA;
    } else {
      for (int i = 0; i < n_ops; ++i) cudaEventElapsedTime(&host_ms[i], ev[i], ev[i + 1]);
    }
  }
  for (int i = 0; i <= n_ops; ++i) cudaEventDestroy(ev[i]);
  delete[] ev;
  return rc;
}

static int run_impl(const VqvsOp* ops, int n_ops, void* stream, cudaEvent_t* ev) {
  if (!ops || n_ops < 0) {
    vqvs::set_error("vqvs_run: bad program");
    return VQVS_EINVAL;
  }
  if (ev) cudaEventRecord(ev[0], (cudaStream_t)stream);
  for (int i = 0; i < n_ops; ++i) {
    int rc = VQVS_OK;
    const void* p = ops[i].desc;
    switch (ops[i].kind) {
      case VQVS_OP_CONV_SIMT: rc = vqvs_conv1d_fused((const VqvsConv*)p, stream); break;
      case VQVS_OP_CONV_UMMA: rc = vqvs_conv1d_umma((const VqvsConv*)p, stream); break;
      case VQVS_OP_GN_FINALIZE: rc = vqvs_gn_finalize((const VqvsGnFinalize*)p, stream); break;
      case VQVS_OP_CONV_IN: rc = vqvs_conv_in((const VqvsConvIn*)p, stream); break;
      case VQVS_OP_CONV_OUT: rc = vqvs_conv_out((const VqvsConvOut*)p, stream); break;
      case VQVS_OP_TIME_EMBED: rc = vqvs_time_embed((const VqvsTimeEmbed*)p, stream); break;
      case VQVS_OP_FILM: rc = vqvs::run_film((const VqvsFilm*)p, stream); break;
      case VQVS_OP_DDPM_FINISH: rc = vqvs_ddpm_finish((const VqvsDdpmFinish*)p, stream); break;
      case VQVS_OP_GN_BWD_PREP: rc = vqvs_gn_bwd_prep(((const VqvsGnBwdPrep*)p)->gn, ((const VqvsGnBwdPrep*)p)->prep, stream); break;
      case VQVS_OP_GELU_BWD: rc = vqvs_gelu_bwd((const VqvsGeluBwd*)p, stream); break;
      case VQVS_OP_GN_BWD_FINALIZE: rc = vqvs_gn_bwd_finalize((const VqvsGnBwdFinalize*)p, stream); break;
      case VQVS_OP_AFFINE3: rc = vqvs_affine3((const VqvsAffine3*)p, stream); break;
      case VQVS_OP_CONV_IN_BWD: rc = vqvs_conv_in_bwd((const VqvsConvInBwd*)p, stream); break;
      case VQVS_OP_ATTNPOOL_FWD: rc = vqvs_attnpool_fwd((const VqvsAttnPool*)p, stream); break;
      case VQVS_OP_ATTNPOOL_BWD: rc = vqvs_attnpool_bwd((const VqvsAttnPool*)p, stream); break;
      case VQVS_OP_CLS_HEAD_FWD: rc = vqvs_cls_head_fwd((const VqvsClsHead*)p, stream); break;
      case VQVS_OP_CLS_HEAD_BWD: rc = vqvs_cls_head_bwd((const VqvsClsHead*)p, stream); break;
      case VQVS_OP_MEMSET: {
        const VqvsMemset* m = (const VqvsMemset*)p;
        cudaError_t e = cudaMemsetAsync(m->ptr, 0, (size_t)m->bytes, (cudaStream_t)stream);
        if (e != cudaSuccess) {
          vqvs::set_error("memset: %s", cudaGetErrorString(e));
          rc = VQVS_ECUDA;
        }
        break;
      }
      default:
        vqvs::set_error("unknown op kind %d", ops[i].kind);
        rc = VQVS_EINVAL;
    }
    if (rc != VQVS_OK) {
      char msg[400];
      strncpy(msg, vqvs::g_err, sizeof(msg) - 1);
      msg[sizeof(msg) - 1] = 0;
      vqvs::set_error("vqvs_run: op %d (kind %d) failed: %s", i, ops[i].kind, msg);
      return rc;
    }
    if (ops[i].kind > 0 && ops[i].kind < 32) __atomic_fetch_add(&vqvs::g_launches[ops[i].kind], 1ull, __ATOMIC_RELAXED);
    if (ev) cudaEventRecord(ev[i + 1], (cudaStream_t)stream);
  }
  return VQVS_OK;
}

extern "C" int64_t vqvs_workspace_bytes(int kind, const void* desc) {
  if (!desc || kind < VQVS_OP_CONV_SIMT || kind > VQVS_OP_CLS_HEAD_BWD) {
    vqvs::set_error("vqvs_workspace_bytes: unknown op kind %d or null descriptor", kind);
    return -1;
  }
  if (kind == VQVS_OP_ATTNPOOL_FWD || kind == VQVS_OP_ATTNPOOL_BWD) {
    const VqvsAttnPool* a = (const VqvsAttnPool*)desc;
    return vqvs_attnpool_workspace_bytes(a->batch, a->c, a->t, a->heads);
  }
  return 0;  // descriptor buffers + shared memory only
}

extern "C" int vqvs_launch_counts(unsigned long long* out16) {  // 32 slots
  if (!out16) {
    vqvs::set_error("vqvs_launch_counts: null pointer");
    return VQVS_EINVAL;
  }
  for (int i = 0; i < 32; ++i) out16[i] = __atomic_load_n(&vqvs::g_launches[i], __ATOMIC_RELAXED);
  return VQVS_OK;
}
