"""ctypes binding of libvqvs.so (include/vqvs.h).  No fallback: if the library is
missing or a call fails, a RuntimeError is raised."""

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# VQVS_LIB selects another build of the same sources (profiling variants); there is still no fallback
LIB_PATH = os.environ.get("VQVS_LIB") or os.path.join(_HERE, "libvqvs.so")

# --- constants (keep in sync with include/vqvs.h) ----------------------------
ABI_VERSION = 7
RESIZE_NONE, RESIZE_DOWN2, RESIZE_UP2 = 0, 1, 2
SKIP_NONE, SKIP_IDENTITY, SKIP_CONV1X1 = 0, 1, 2
OUT_EPS, OUT_PREV, OUT_X0_SUM = 0, 1, 2
OP_CONV_SIMT, OP_CONV_UMMA, OP_GN_FINALIZE, OP_CONV_IN, OP_CONV_OUT = 1, 2, 3, 4, 5
OP_TIME_EMBED, OP_FILM, OP_MEMSET, OP_DDPM_FINISH = 6, 7, 8, 9
OP_GN_BWD_PREP, OP_GELU_BWD, OP_GN_BWD_FINALIZE, OP_AFFINE3, OP_CONV_IN_BWD = 10, 11, 12, 13, 14
OP_ATTNPOOL_FWD, OP_ATTNPOOL_BWD, OP_CLS_HEAD_FWD, OP_CLS_HEAD_BWD = 15, 16, 17, 18
CONV_PAIR_STATS = 1024  # VqvsConv.reserved_ flag
CONV_STAT_GRAN_SHIFT = 12  # bits 12..15 of VqvsConv.reserved_: log2 of the statistics granularity
CONV_PREC_SHIFT = 16  # bits 16..17 of VqvsConv.reserved_: tensor-core operand format
PREC_BF16X3, PREC_F16 = 0, 1

_i32, _i64, _p = C.c_int32, C.c_int64, C.c_void_p


class Conv(C.Structure):
    _fields_ = [
        ("batch", _i32), ("c_a", _i32), ("c_b", _i32), ("t_in", _i32), ("c_out", _i32), ("t_out", _i32),
        ("ksize", _i32), ("dilation", _i32), ("resize", _i32), ("act", _i32), ("skip_mode", _i32),
        ("s_a", _i32), ("s_b", _i32), ("t_skip", _i32), ("skip_resize", _i32), ("reserved_", _i32),
        ("xa", _p), ("xb", _p), ("scale", _p), ("shift", _p), ("w", _p), ("bias", _p),
        ("sa", _p), ("sb", _p), ("w_skip", _p), ("b_skip", _p), ("w_packed", _p), ("out", _p), ("stats_out", _p),
        ("gn", _p),
    ]


class GnFinalize(C.Structure):
    _fields_ = [
        ("batch", _i32), ("c_a", _i32), ("c_b", _i32), ("groups", _i32), ("count", _i64),
        ("stats_a", _p), ("stats_b", _p), ("gamma", _p), ("beta", _p), ("film", _p), ("film_stride", _i64),
        ("scale", _p), ("shift", _p),
    ]


class ConvIn(C.Structure):
    _fields_ = [
        ("batch", _i32), ("c_out", _i32), ("t", _i32), ("t_cond", _i32),
        ("x", _p), ("w", _p), ("bias", _p), ("cond", _p), ("out", _p), ("stats_out", _p),
    ]


class ConvOut(C.Structure):
    _fields_ = [
        ("batch", _i32), ("c_in", _i32), ("t", _i32), ("mode", _i32),
        ("x", _p), ("scale", _p), ("shift", _p), ("w", _p), ("bias", _p), ("x_t", _p), ("noise", _p),
        ("coef", _p), ("out", _p), ("x0_sum", _p),
    ]


class DdpmFinish(C.Structure):
    _fields_ = [
        ("batch", _i32), ("t", _i32), ("use_x0_mean", _i32),
        ("x_t", _p), ("eps", _p), ("noise", _p), ("coef", _p), ("x0_sum", _p), ("out", _p),
    ]


class TimeEmbed(C.Structure):
    _fields_ = [
        ("batch", _i32), ("dim", _i32),
        ("ts", _p), ("freqs", _p), ("w1", _p), ("b1", _p), ("w2", _p), ("b2", _p),
        ("class_embed", _p), ("labels", _p), ("emb", _p), ("gelu_emb", _p),
    ]


class Film(C.Structure):
    _fields_ = [
        ("gelu_emb", _p), ("w_cat", _p), ("b_cat", _p), ("batch", _i32), ("dim", _i32), ("n_out", _i32), ("ab", _p),
    ]


class GnBwdPrep(C.Structure):
    _fields_ = [("gn", _p), ("prep", _p)]


class GeluBwd(C.Structure):
    _fields_ = [("batch", _i32), ("c", _i32), ("t", _i32), ("up", _i32), ("c_total", _i32), ("c_off", _i32),
                ("d_in", _p), ("z", _p), ("prep", _p), ("q", _p), ("acc", _p)]


class GnBwdFinalize(C.Structure):
    _fields_ = [("batch", _i32), ("c", _i32), ("groups", _i32), ("pad_", _i32), ("count", _i64),
                ("acc", _p), ("prep", _p), ("coef", _p)]


class Affine3(C.Structure):
    _fields_ = [("batch", _i32), ("c", _i32), ("t", _i32), ("add_mode", _i32), ("c_total", _i32), ("c_off", _i32),
                ("q", _p), ("z", _p), ("coef", _p), ("add", _p), ("add2", _p), ("out", _p)]


class ConvInBwd(C.Structure):
    _fields_ = [("batch", _i32), ("c", _i32), ("t", _i32), ("pad_", _i32), ("dh", _p), ("w", _p), ("dx", _p)]


class AttnPool(C.Structure):
    _fields_ = [("batch", _i32), ("c", _i32), ("t", _i32), ("heads", _i32), ("c_out", _i32), ("pad_", _i32),
                ("h", _p), ("prep", _p), ("w_qkv", _p), ("b_qkv", _p), ("w_proj", _p), ("b_proj", _p),
                ("ws", _p), ("out", _p), ("d_out", _p), ("d_act", _p)]


class ClsHead(C.Structure):
    _fields_ = [("batch", _i32), ("dim", _i32), ("labels", _i32), ("pad_", _i32),
                ("stem", _p), ("w", _p), ("b", _p), ("logits", _p), ("d_logits", _p), ("d_stem", _p)]


class Mfcc(C.Structure):
    _fields_ = [(n, _i32) for n in "batch t n_fft hop n_bins n_mels n_mfcc frames c_pad ulaw".split()] + \
               [(n, _p) for n in "x window cos_t sin_t fb dct mfcc out".split()]


class Memset(C.Structure):
    _fields_ = [("ptr", _p), ("bytes", _i64)]


class Op(C.Structure):
    _fields_ = [("kind", _i32), ("desc", _p)]


# name -> (restype, argtypes); every symbol declared in include/vqvs.h
SIGNATURES = {
    "vqvs_abi_version": (C.c_int, []),
    "vqvs_last_error": (C.c_char_p, []),
    "vqvs_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vqvs_conv1d_fused": (C.c_int, [C.POINTER(Conv), _p]),
    "vqvs_conv1d_umma": (C.c_int, [C.POINTER(Conv), _p]),
    "vqvs_conv1d_umma_supported": (C.c_int, [C.POINTER(Conv)]),
    "vqvs_packed_weight_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vqvs_pack_conv_weights": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "vqvs_gn_finalize": (C.c_int, [C.POINTER(GnFinalize), _p]),
    "vqvs_channel_stats": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "vqvs_conv_in": (C.c_int, [C.POINTER(ConvIn), _p]),
    "vqvs_conv_out": (C.c_int, [C.POINTER(ConvOut), _p]),
    "vqvs_ddpm_finish": (C.c_int, [C.POINTER(DdpmFinish), _p]),
    "vqvs_ddpm_x0_sum": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, _p, _p]),
    "vqvs_time_embed": (C.c_int, [C.POINTER(TimeEmbed), _p]),
    "vqvs_gelu": (C.c_int, [_p, _p, C.c_int64, _p]),
    "vqvs_film_linear": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "vqvs_vq_argmin": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "vqvs_vq_embed": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "vqvs_gn_bwd_prep": (C.c_int, [C.POINTER(GnFinalize), _p, _p]),
    "vqvs_gelu_bwd": (C.c_int, [C.POINTER(GeluBwd), _p]),
    "vqvs_gn_bwd_finalize": (C.c_int, [C.POINTER(GnBwdFinalize), _p]),
    "vqvs_affine3": (C.c_int, [C.POINTER(Affine3), _p]),
    "vqvs_conv_in_bwd": (C.c_int, [C.POINTER(ConvInBwd), _p]),
    "vqvs_stride_sample": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "vqvs_attnpool_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vqvs_workspace_bytes": (C.c_int64, [C.c_int, C.c_void_p]),
    "vqvs_attnpool_fwd": (C.c_int, [C.POINTER(AttnPool), _p]),
    "vqvs_attnpool_bwd": (C.c_int, [C.POINTER(AttnPool), _p]),
    "vqvs_cls_head_fwd": (C.c_int, [C.POINTER(ClsHead), _p]),
    "vqvs_cls_head_bwd": (C.c_int, [C.POINTER(ClsHead), _p]),
    "vqvs_mfcc39": (C.c_int, [C.POINTER(Mfcc), _p]),
    "vqvs_gelu_add": (C.c_int, [_p, C.c_int, _p, C.c_int, _p, C.c_int, C.c_int, C.c_int, _p]),
    "vqvs_deinterleave2": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, C.c_int, _p]),
    "vqvs_keyed_normal": (C.c_int, [_p, C.c_int, C.c_int64, C.c_uint64, C.c_int64, C.c_int32, _p]),
    "vqvs_run": (C.c_int, [C.POINTER(Op), C.c_int, _p]),
    "vqvs_run_timed": (C.c_int, [C.POINTER(Op), C.c_int, _p, C.POINTER(C.c_float)]),
    "vqvs_launch_counts": (C.c_int, [C.POINTER(C.c_uint64)]),
    "vqvs_debug_prof": (C.c_int, [C.POINTER(C.c_uint64)]),
    "vqvs_debug_geo": (C.c_int, [C.POINTER(Conv), C.POINTER(C.c_int)]),
    "vqvs_umma_selftest": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """dlopen libvqvs.so and bind every declared symbol (no GPU needed)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "vq_voice_swap_b200 has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if lib.vqvs_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libvqvs ABI {lib.vqvs_abi_version()} != binding {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def check(rc: int, what: str = "libvqvs") -> None:
    if rc != 0:
        msg = load().vqvs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


OP_NAMES = {OP_CONV_SIMT: "conv_simt", OP_CONV_UMMA: "conv_umma", OP_GN_FINALIZE: "gn_finalize", OP_CONV_IN: "conv_in",
            OP_CONV_OUT: "conv_out", OP_TIME_EMBED: "time_embed", OP_FILM: "film", OP_MEMSET: "memset",
            OP_DDPM_FINISH: "ddpm_finish", OP_GN_BWD_PREP: "gn_bwd_prep", OP_GELU_BWD: "gelu_bwd",
            OP_GN_BWD_FINALIZE: "gn_bwd_finalize", OP_AFFINE3: "affine3", OP_CONV_IN_BWD: "conv_in_bwd",
            OP_ATTNPOOL_FWD: "attnpool_fwd", OP_ATTNPOOL_BWD: "attnpool_bwd", OP_CLS_HEAD_FWD: "cls_head_fwd",
            OP_CLS_HEAD_BWD: "cls_head_bwd"}


def launch_counts() -> dict:
    """Ops executed through vqvs_run since the library was loaded, by name (include/vqvs.h: vqvs_launch_counts)."""
    buf = (C.c_uint64 * 32)()
    check(load().vqvs_launch_counts(buf), "vqvs_launch_counts")
    return {name: int(buf[kind]) for kind, name in OP_NAMES.items()}


def stream_ptr(device=None) -> int:
    """torch's current CUDA stream on `device` (default: the current device) as a cudaStream_t."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
