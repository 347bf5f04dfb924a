"""Host-side program builder: turns a UNet (predictor / encoder / single ResBlock) plus a
(batch, length) into a static list of libvqvs ops over preallocated HBM buffers.

HBM layout per plan (all fp32, NCT, dense):
  * skip stack: one buffer per down-path tensor (reference models/unet.py:141-147) -- they
    stay resident until the up path consumes them; the concat of unet.py:156 is never built,
    consumers read the two sources through two pointers;
  * two ping-pong buffers for middle/up outputs and one scratch for the intra-block tensor h1;
  * a float64 statistics arena: (sum, sumsq) per (sample, channel) of every tensor that feeds a
    GroupNorm, filled by the producing kernel's epilogue and zeroed once per forward;
  * per-(sample, channel) GroupNorm/FiLM scale+shift scratch, the FiLM table ab[batch, sum 2*C_out].
Weights are re-laid once per parameter version into the tcgen05 operand image (bf16 hi/lo).
"""

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import List, Optional, Sequence

import torch

from . import lib as L


def backend_default() -> str:
    """'umma' (tcgen05) unless VQVS_BACKEND=simt forces the CUDA-core kernels."""
    return os.environ.get("VQVS_BACKEND", "umma")


def resize_mode(scale_factor: float) -> int:
    if scale_factor == 1.0:
        return L.RESIZE_NONE
    if scale_factor == 0.5:
        return L.RESIZE_DOWN2
    if scale_factor == 2.0:
        return L.RESIZE_UP2
    raise ValueError(f"unsupported scale factor {scale_factor}: the CUDA path implements 0.5, 1 and 2")


def _resized(t: int, mode: int) -> int:
    return t // 2 if mode == L.RESIZE_DOWN2 else t * 2 if mode == L.RESIZE_UP2 else t


class PlanCache:
    """Tiny LRU of compiled plans keyed by (batch, length, ...)."""

    def __init__(self, capacity: int = 2):
        self.capacity = capacity
        self.items = OrderedDict()
        self.signature = None

    def clear(self):
        self.items.clear()

    def get(self, key, signature, build):
        if signature != self.signature:
            self.items.clear()
            self.signature = signature
        plan = self.items.get(key)
        if plan is None:
            plan = build()
            self.items[key] = plan
            while len(self.items) > self.capacity:
                self.items.popitem(last=False)
        else:
            self.items.move_to_end(key)
        return plan

    def __deepcopy__(self, memo):
        return PlanCache(self.capacity)

    def __getstate__(self):
        return {"capacity": self.capacity}

    def __setstate__(self, state):
        self.__init__(state["capacity"])


def _signature(module: torch.nn.Module):
    """(pointer, version) of every parameter: plans and packed weight images are rebuilt when it changes.  In-place
    updates that bypass autograd's version counter (`p.data.mul_()`) are invisible here: call invalidate(module)."""
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def invalidate(module: torch.nn.Module) -> None:
    """Drop every compiled plan and derived weight image of `module` (after editing parameters through `.data`)."""
    for m in module.modules():
        plans = getattr(m, "_plans", None)
        if isinstance(plans, PlanCache):
            plans.clear()
            plans.signature = None
        if hasattr(m, "_vqvs_weights"):
            m._vqvs_weights = None


def _check_module(net: torch.nn.Module) -> None:
    """Plans embed raw parameter pointers that the kernels read as dense fp32: refuse anything else loudly."""
    for name, p in net.named_parameters():
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise TypeError(f"parameter {name} is {p.dtype}{'' if p.is_contiguous() else ', non-contiguous'}: the sm_100a path "
                            "needs contiguous float32 parameters (no .half()/.double() models)")
    if net.training and any(getattr(m, "dropout", 0.0) for m in net.modules()):
        raise RuntimeError("dropout > 0 in train mode is not implemented on the sm_100a path (sampling only): call .eval()")


def _require_cuda(*tensors):
    L.load()
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "vq_voice_swap_b200 executes on CUDA (sm_100a) only and has no CPU fallback; "
                "move the model and its inputs to a CUDA device"
            )


def _f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Act:
    """An activation tensor [batch, c, t] with its statistics slot."""

    def __init__(self, buf: torch.Tensor, c: int, t: int, stats: Optional[torch.Tensor]):
        self.buf, self.c, self.t, self.stats = buf, c, t, stats
        self.gran = 0         # gcd of the GroupNorm group sizes (and concat offsets) that consume the statistics
        self.producer = None  # conv descriptor that writes this tensor

    @property
    def ptr(self):
        return self.buf.data_ptr()

    @property
    def stats_ptr(self):
        return 0 if self.stats is None else self.stats.data_ptr()


class Plan:
    """A compiled launch program and the buffers it points into."""

    def __init__(self, device, batch: int, backend: str):
        self.device, self.batch, self.backend = device, batch, backend
        self.descs = []     # (kind, ctypes struct)
        self.keep = []      # tensors referenced by raw pointer
        self.ops = None
        self.n_launch = 0   # kernels launched per run (memsets excluded)
        self.slots = {}     # named structs patched per call
        self.produced = []  # activations written by conv ops (statistics granularity resolved in compile())

    # -- buffers -------------------------------------------------------------
    def empty(self, *shape, dtype=torch.float32):
        t = torch.empty(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def hold(self, t):
        self.keep.append(t)
        return t

    # -- ops -----------------------------------------------------------------
    def add(self, kind, desc, name=None):
        self.descs.append((kind, desc))
        if kind != L.OP_MEMSET:
            self.n_launch += 1
        if name:
            self.slots[name] = desc
        return desc

    def compile(self):
        # statistics granularity flags: known only now that every consumer GroupNorm has been emitted
        for act in self.produced:
            gran = act.gran & -act.gran if act.gran else 1  # power-of-two part
            if act.stats is not None and gran >= 2:
                act.producer.reserved_ |= L.CONV_PAIR_STATS | (min(gran, 16).bit_length() - 1) << L.CONV_STAT_GRAN_SHIFT
        arr = (L.Op * len(self.descs))()
        for i, (kind, desc) in enumerate(self.descs):
            arr[i].kind = kind
            arr[i].desc = C.addressof(desc)
        self.ops = arr
        return self

    def run(self):
        # the launches go to the plan's device and to torch's current stream ON THAT DEVICE, whatever device is current
        with torch.cuda.device(self.device):
            L.check(L.load().vqvs_run(self.ops, len(self.descs), L.stream_ptr(self.device)), "vqvs_run")


class Weights:
    """Device-side weight images derived from a module's parameters (packed once per version)."""

    def __init__(self):
        self.packed = {}
        self.film_w = self.film_b = None
        self.film_offsets = {}
        self.film_total = 0
        self.freqs = None
        self.signature = None


class Packed:
    """tcgen05 operand image of one conv (+ its 1x1 skip) and the operand format it was built for."""

    def __init__(self, img: torch.Tensor, prec: int):
        self.img, self.prec = img, prec


def conv_precision(c_out: int, base_channels: Optional[int], rel_length: float = 1.0) -> int:
    """Operand format of a predictor conv (include/vqvs.h VQVS_PREC_*).

    One fp16 product per tap where it is cheap in error: the deep levels (C_out >= 4 * base_channels: 256 and 512 channels in
    unet64, bound by the tensor pipe under bf16x3) and the 2 * base_channels blocks whose output is at most 1/16 of the input
    length (128 channels at T = 4000 in unet64); everything else keeps the bf16 hi/lo split.  Measured on the B200
    (tools/precision_policy_study.py, unet64 forward at T = 64000 against the fp32 oracle): 1.2e-5 with bf16x3 everywhere,
    4.9e-5 with the first rule, 8.9e-5 with both -- the budget is 2e-4, a fifth of the 1e-3 the north star allows.  Wider
    choices are opt-in: VQVS_F16_FROM=2 (all 2 * bc blocks: 3.7e-4) and =1 (everything: 1.0e-3 per forward, 2.8e-4 on a
    50-step sample -- the TF32 class of the reference's own GPU path); VQVS_PREC=bf16x3 | f16 forces one format everywhere."""
    forced = os.environ.get("VQVS_PREC")
    if forced:
        return {"bf16x3": L.PREC_BF16X3, "f16": L.PREC_F16}[forced]
    if base_channels is None:
        return L.PREC_BF16X3
    if c_out >= int(os.environ.get("VQVS_F16_FROM", "4")) * base_channels:
        return L.PREC_F16
    if c_out >= 2 * base_channels and rel_length <= 1.0 / 16 and os.environ.get("VQVS_F16_SHORT", "1") == "1":
        return L.PREC_F16
    return L.PREC_BF16X3


def pack_weights(weight: torch.Tensor, skip_weight: Optional[torch.Tensor], prec: int = L.PREC_BF16X3) -> Optional[Packed]:
    """fp32 [c_out, c_in, k] (+ optional 1x1 skip [c_out, c_skip, 1]) -> tcgen05 operand image, or None if unsupported."""
    lib = L.load()
    c_out, c_in, k = weight.shape
    c_skip = skip_weight.shape[1] if skip_weight is not None else 0
    nbytes = lib.vqvs_packed_weight_bytes(c_out, c_in, k, c_skip, prec)
    if nbytes <= 0:
        return None
    dev = weight.device
    img = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    w = _f32(weight)
    ws = _f32(skip_weight) if skip_weight is not None else None
    with torch.cuda.device(dev):
        L.check(lib.vqvs_pack_conv_weights(L.ptr(w), L.ptr(ws), c_out, c_in, k, c_skip, prec, L.ptr(img), L.stream_ptr(dev)),
                "vqvs_pack_conv_weights")
        torch.cuda.current_stream(dev).synchronize()  # w / ws temporaries may be freed after this
    return Packed(img, prec)


def _pack(conv: torch.nn.Conv1d, skip: Optional[torch.nn.Conv1d], prec: int = L.PREC_BF16X3) -> Optional[Packed]:
    return pack_weights(conv.weight, skip.weight if skip is not None else None, prec)


def _skip_proj(block):
    proj = block.skip[1]
    return proj if isinstance(proj, torch.nn.Conv1d) else None


def _tail_conv(block):
    return block.post_cond[len(block.post_cond) - 1]


def weights_for(net, blocks: Sequence, backend: str, reduced_precision: bool = False) -> Weights:
    """(Re)build the derived weight images of `net` when any parameter changed.

    reduced_precision: apply conv_precision()'s per-level rule (the predictor and the guidance stem; the VQ-VAE encoder
    keeps bf16x3 everywhere because its outputs decide code indices)."""
    sig = _signature(net)
    w = getattr(net, "_vqvs_weights", None)
    key = (sig, backend, reduced_precision) + tuple(os.environ.get(k) for k in ("VQVS_PREC", "VQVS_F16_FROM", "VQVS_BF16X3_FIRST", "VQVS_BF16X3_LAST", "VQVS_F16_BLOCKS", "VQVS_F16_SHORT"))
    if w is not None and w.signature == key:
        return w
    w = Weights()
    w.signature = key
    device = next(net.parameters()).device
    bc = getattr(net, "base_channels", None) if reduced_precision else None
    with torch.no_grad():
        if backend == "umma":
            extra_f16 = set()  # VQVS_F16_BLOCKS="6-14,46-52": block indices (down + middle + up order) forced to fp16 (study knob)
            for part in filter(None, os.environ.get("VQVS_F16_BLOCKS", "").split(",")):
                lo, _, hi = part.partition("-")
                extra_f16.update(range(int(lo), int(hi or lo) + 1))
            keep_first = int(os.environ.get("VQVS_BF16X3_FIRST", "0"))  # (study knobs: blocks at either end of the network
            keep_last = int(os.environ.get("VQVS_BF16X3_LAST", "0"))    #  that keep the bf16 hi/lo split whatever the rule says)
            rel = 1.0  # length of the block's output relative to the network input (both convs of a block run at it)
            for bi, blk in enumerate(blocks):
                rel *= float(getattr(blk, "scale_factor", 1.0) or 1.0)
                prec = conv_precision(blk.out_channels, bc, rel)
                if bc is not None and (bi < keep_first or bi >= len(blocks) - keep_last) and not os.environ.get("VQVS_PREC"):
                    prec = L.PREC_BF16X3
                if bc is not None and bi in extra_f16:
                    prec = L.PREC_F16
                w.packed[(id(blk), 1)] = _pack(blk.pre_cond[2], None, prec)
                w.packed[(id(blk), 2)] = _pack(_tail_conv(blk), _skip_proj(blk), prec)
            for name in ("cond_proj",):
                conv = getattr(net, name, None)
                if isinstance(conv, torch.nn.Conv1d):
                    w.packed[(id(net), name)] = _pack(conv, None)
            head = getattr(net, "out", None)
            if head is not None and isinstance(head[1], torch.nn.Conv1d) and head[1].weight.shape[0] > 1:
                w.packed[(id(net), "out")] = _pack(head[1], None)
        film = [b for b in blocks if getattr(b, "emb_channels", None)]
        if film:
            off = 0
            for b in film:
                w.film_offsets[id(b)] = off
                off += 2 * b.out_channels
            w.film_total = off
            w.film_w = torch.cat([_f32(b.cond_layers[1].weight) for b in film], dim=0).contiguous()
            w.film_b = torch.cat([_f32(b.cond_layers[1].bias) for b in film], dim=0).contiguous()
        if hasattr(net, "time_embed"):
            dim = net.time_embed.channels
            half = dim // 2
            # built exactly like reference models/wavegrad.py:363-370 (fp32 on the host, then moved)
            w.freqs = (
                torch.exp(-math.log(100.0 / 0.1) * torch.arange(start=0, end=half, dtype=torch.float32) / (half - 1))
                * 100.0
            ).to(device)
    net._vqvs_weights = w
    return w


# ---------------------------------------------------------------------------------------------
# op emitters
# ---------------------------------------------------------------------------------------------
def _emit_gn(plan: Plan, srcs: List[Act], gn: torch.nn.GroupNorm, scale, shift, film_ptr=0, film_stride=0,
             standalone: bool = True):
    """GroupNorm(+FiLM) finalize.  standalone=False only builds the descriptor (for a conv that fuses it)."""
    d = L.GnFinalize()
    d.batch = plan.batch
    d.c_a = srcs[0].c
    d.c_b = srcs[1].c if len(srcs) > 1 else 0
    d.groups = gn.num_groups
    d.count = srcs[0].t
    d.stats_a = srcs[0].stats_ptr
    d.stats_b = srcs[1].stats_ptr if len(srcs) > 1 else 0
    d.gamma, d.beta = L.ptr(gn.weight), L.ptr(gn.bias)
    # statistics granularity each producer may use: groups of the concatenated tensor must be unions of whole
    # producer granules, so a source starting at channel offset `off` can merge gcd(group size, off) channels
    gs, off = (d.c_a + d.c_b) // gn.num_groups, 0
    for src in srcs:
        src.gran = math.gcd(src.gran, math.gcd(gs, off))
        off += src.c
    d.film, d.film_stride = film_ptr, film_stride
    d.scale, d.shift = L.ptr(scale), L.ptr(shift)
    if standalone:
        plan.add(L.OP_GN_FINALIZE, d)
    else:
        plan.keep.append(d)
    return d


def _emit_conv(plan: Plan, srcs: List[Act], conv: torch.nn.Conv1d, out: Act, *, scale=None, shift=None,
               resize=L.RESIZE_NONE, skip_srcs: Optional[List[Act]] = None, skip_proj=None, skip_resize=L.RESIZE_NONE,
               packed=None, name=None, norm=None):
    """norm = (GroupNorm module, film_ptr, film_stride): the GroupNorm(+FiLM) in front of this conv.  On the tcgen05
    path its finalize is fused into the conv kernel (VqvsConv.gn); otherwise a vqvs_gn_finalize op is emitted first."""
    d = L.Conv()
    d.batch = plan.batch
    d.c_a, d.c_b = srcs[0].c, (srcs[1].c if len(srcs) > 1 else 0)
    d.t_in = srcs[0].t
    d.c_out, d.t_out = out.c, out.t
    d.ksize, d.dilation = conv.kernel_size[0], conv.dilation[0]
    d.resize = resize
    d.act = 1 if scale is not None else 0
    d.xa, d.xb = srcs[0].ptr, (srcs[1].ptr if len(srcs) > 1 else 0)
    d.scale, d.shift = L.ptr(scale), L.ptr(shift)
    d.w, d.bias = L.ptr(conv.weight), L.ptr(conv.bias)
    if skip_srcs:
        d.skip_mode = L.SKIP_CONV1X1 if skip_proj is not None else L.SKIP_IDENTITY
        d.s_a, d.s_b = skip_srcs[0].c, (skip_srcs[1].c if len(skip_srcs) > 1 else 0)
        d.t_skip = skip_srcs[0].t
        d.skip_resize = skip_resize
        d.sa, d.sb = skip_srcs[0].ptr, (skip_srcs[1].ptr if len(skip_srcs) > 1 else 0)
        if skip_proj is not None:
            d.w_skip, d.b_skip = L.ptr(skip_proj.weight), L.ptr(skip_proj.bias)
    d.w_packed = L.ptr(packed.img) if packed is not None else 0
    d.reserved_ = int(os.environ.get("VQVS_DEBUG_FLAGS", "0"))  # kernel ablation switches (profiling only)
    if packed is not None:
        d.reserved_ |= packed.prec << L.CONV_PREC_SHIFT
    d.out, d.stats_out = out.ptr, out.stats_ptr
    out.producer = d
    plan.produced.append(out)
    kind = L.OP_CONV_SIMT
    if plan.backend == "umma" and packed is not None and L.load().vqvs_conv1d_umma_supported(C.byref(d)):
        kind = L.OP_CONV_UMMA
    if norm is not None:
        gn, film_ptr, film_stride = norm
        fuse = kind == L.OP_CONV_UMMA and not os.environ.get("VQVS_NO_GN_FUSION")
        fin = _emit_gn(plan, srcs, gn, scale, shift, film_ptr, film_stride, standalone=not fuse)
        if fuse:
            d.gn = C.addressof(fin)
    plan.add(kind, d, name)
    return kind


def _emit_block(plan: Plan, blk, srcs: List[Act], h1: Act, out: Act, w: Weights, scratch, ab: Optional[torch.Tensor]):
    """One reference ResBlock = GN finalize, fused conv1, GN(+FiLM) finalize, fused conv2(+skip)."""
    sc_a, sh_a, sc_b, sh_b = scratch
    mode = resize_mode(blk.scale_factor)
    _emit_conv(plan, srcs, blk.pre_cond[2], h1, scale=sc_a, shift=sh_a, resize=mode, packed=w.packed.get((id(blk), 1)),
               norm=(blk.pre_cond[0][0], 0, 0))
    film_ptr = film_stride = 0
    if getattr(blk, "emb_channels", None):
        film_ptr = ab.data_ptr() + 4 * w.film_offsets[id(blk)]
        film_stride = w.film_total
    _emit_conv(plan, [h1], _tail_conv(blk), out, scale=sc_b, shift=sh_b, skip_srcs=srcs, skip_proj=_skip_proj(blk),
               skip_resize=mode, packed=w.packed.get((id(blk), 2)), norm=(blk.pre_cond[3], film_ptr, film_stride))


class _Alloc:
    """Statistics arena + activation buffers for one plan."""

    def __init__(self, plan: Plan, stats_elems: int):
        self.plan = plan
        self.arena = plan.empty(max(stats_elems, 2), dtype=torch.float64)
        self.used = 0

    def stats(self, c: int):
        n = self.plan.batch * c * 2
        view = self.arena[self.used:self.used + n]
        self.used += n
        assert self.used <= self.arena.numel()
        return view

    def act(self, c: int, t: int, buf: Optional[torch.Tensor] = None, stats: bool = True) -> Act:
        if buf is None:
            buf = self.plan.empty(self.plan.batch, c, t)
        else:
            buf = buf[: self.plan.batch * c * t].view(self.plan.batch, c, t)
        return Act(buf, c, t, self.stats(c) if stats else None)


def _scratch(plan: Plan, c_max: int):
    return tuple(plan.empty(plan.batch, c_max) for _ in range(4))


def _memset_op(plan: Plan, tensor: torch.Tensor):
    m = L.Memset()
    m.ptr, m.bytes = tensor.data_ptr(), tensor.numel() * tensor.element_size()
    plan.add(L.OP_MEMSET, m)


# ---------------------------------------------------------------------------------------------
# UNetPredictor
# ---------------------------------------------------------------------------------------------
def _predictor_blocks(net):
    return list(net.down_blocks) + list(net.middle_blocks) + list(net.up_blocks)


def build_predictor_plan(net, batch: int, t: int, t_cond: Optional[int], backend: str, keep_activations: bool = False) -> Plan:
    """keep_activations: every block gets its own intra-block and output buffers and the plan records
    `plan.saved = [(block, [source Acts], u, out, resize mode)]` plus `plan.head` -- what a backward program needs
    (guidance.PredictorGuidancePlans); sampling plans reuse three buffers instead."""
    rate = net.downsample_rate
    if t % rate:
        raise ValueError(f"sequence length {t} must be divisible by the UNet downsample rate {rate}")
    if net.in_channels != 1:
        raise ValueError("the CUDA predictor path implements in_channels == 1 (waveforms)")
    device = next(net.parameters()).device
    _check_module(net)
    blocks = _predictor_blocks(net)
    w = weights_for(net, blocks, backend, reduced_precision=not keep_activations)
    plan = Plan(device, batch, backend)
    plan.weights = w
    plan.saved = []
    bc = net.base_channels
    emb_dim = 4 * bc

    c_all = [bc] + [b.out_channels for b in blocks] * 2
    alloc = _Alloc(plan, 2 * batch * sum(c_all))
    c_max = max(b.channels for b in blocks)
    scratch = _scratch(plan, c_max)
    max_elems = batch * max(
        max(b.out_channels * _resized(t_l, resize_mode(b.scale_factor)) for b, t_l in _walk_lengths(net, t)),
        bc * t,
    )
    ping = [plan.empty(max_elems), plan.empty(max_elems)] if not keep_activations else None
    h1_buf = plan.empty(max_elems) if not keep_activations else None

    # inputs that change per call are patched into these structs
    plan.ts = plan.empty(batch)
    plan.labels = plan.empty(batch, dtype=torch.int64) if net.num_labels is not None else None
    plan.emb = plan.empty(batch, emb_dim)
    plan.gelu_emb = plan.empty(batch, emb_dim)
    plan.ab = plan.empty(batch, w.film_total)
    plan.coef = plan.empty(batch, 8)
    plan.x0_sum = plan.empty(batch, dtype=torch.float64)
    plan.eps = plan.empty(batch, net.out_channels, t)

    _memset_op(plan, alloc.arena)
    te = L.TimeEmbed()
    te.batch, te.dim = batch, emb_dim
    te.ts, te.freqs = plan.ts.data_ptr(), w.freqs.data_ptr()
    te.w1, te.b1 = L.ptr(net.time_embed.proj.weight), L.ptr(net.time_embed.proj.bias)
    te.w2, te.b2 = L.ptr(net.time_embed_extra[1].weight), L.ptr(net.time_embed_extra[1].bias)
    if net.num_labels is not None:
        te.class_embed, te.labels = L.ptr(net.class_embed.weight), plan.labels.data_ptr()
    te.emb, te.gelu_emb = plan.emb.data_ptr(), plan.gelu_emb.data_ptr()
    plan.add(L.OP_TIME_EMBED, te)
    fl = L.Film()
    fl.gelu_emb, fl.w_cat, fl.b_cat = plan.gelu_emb.data_ptr(), w.film_w.data_ptr(), w.film_b.data_ptr()
    fl.batch, fl.dim, fl.n_out, fl.ab = batch, emb_dim, w.film_total, plan.ab.data_ptr()
    plan.add(L.OP_FILM, fl)

    cond_out = None
    if net.cond_channels is not None:
        if t_cond is None:
            raise ValueError("conditional predictor needs a cond sequence")
        cond_in = Act(plan.empty(batch, net.cond_channels, t_cond), net.cond_channels, t_cond, None)
        plan.cond_in = cond_in.buf
        cond_out = Act(plan.empty(batch, bc, t_cond), bc, t_cond, None)
        _emit_conv(plan, [cond_in], net.cond_proj, cond_out, packed=w.packed.get((id(net), "cond_proj")))

    h = alloc.act(bc, t)
    ci = L.ConvIn()
    ci.batch, ci.c_out, ci.t = batch, bc, t
    ci.t_cond = t_cond or 0
    ci.w, ci.bias = L.ptr(net.in_conv.weight), L.ptr(net.in_conv.bias)
    ci.cond = cond_out.ptr if cond_out is not None else 0
    ci.out, ci.stats_out = h.ptr, h.stats_ptr
    plan.add(L.OP_CONV_IN, ci, "conv_in")

    plan.h0 = h
    skips = [h]
    cur_t = t
    own = keep_activations  # True: no buffer is ever reused
    for blk in net.down_blocks:
        mode = resize_mode(blk.scale_factor)
        t_out = _resized(cur_t, mode)
        h1 = alloc.act(blk.out_channels, t_out, None if own else h1_buf)
        out = alloc.act(blk.out_channels, t_out)
        _emit_block(plan, blk, [h], h1, out, w, scratch, plan.ab)
        plan.saved.append((blk, [h], h1, out, mode))
        h, cur_t = out, t_out
        skips.append(h)
    flip = 0
    for blk in net.middle_blocks:
        h1 = alloc.act(blk.out_channels, cur_t, None if own else h1_buf)
        out = alloc.act(blk.out_channels, cur_t, None if own else ping[flip])
        _emit_block(plan, blk, [h], h1, out, w, scratch, plan.ab)
        plan.saved.append((blk, [h], h1, out, L.RESIZE_NONE))
        h, flip = out, flip ^ 1
    period = net.depth_mult + 2
    for i, blk in enumerate(net.up_blocks):
        srcs = [h] if i % period == period - 1 else [h, skips.pop()]
        mode = resize_mode(blk.scale_factor)
        t_out = _resized(cur_t, mode)
        h1 = alloc.act(blk.out_channels, t_out, None if own else h1_buf)
        out = alloc.act(blk.out_channels, t_out, None if own else ping[flip])
        _emit_block(plan, blk, srcs, h1, out, w, scratch, plan.ab)
        plan.saved.append((blk, srcs, h1, out, mode))
        h, cur_t, flip = out, t_out, flip ^ 1
    plan.h_last = h
    if not own:
        plan.saved = []

    head_gn, head_conv = net.out[0][0], net.out[1]
    _emit_gn(plan, [h], head_gn, scratch[0], scratch[1])
    if net.out_channels == 1:
        co = L.ConvOut()
        co.batch, co.c_in, co.t, co.mode = batch, bc, t, L.OUT_EPS
        co.x, co.scale, co.shift = h.ptr, scratch[0].data_ptr(), scratch[1].data_ptr()
        co.w, co.bias = L.ptr(head_conv.weight), L.ptr(head_conv.bias)
        co.coef, co.x0_sum, co.out = plan.coef.data_ptr(), plan.x0_sum.data_ptr(), plan.eps.data_ptr()
        plan.add(L.OP_CONV_OUT, co, "conv_out")
    else:
        out = Act(plan.eps, net.out_channels, t, None)
        _emit_conv(plan, [h], head_conv, out, scale=scratch[0], shift=scratch[1], packed=w.packed.get((id(net), "out")))
    return plan.compile()


def _walk_lengths(net, t: int):
    cur = t
    for blk in _predictor_blocks(net):
        yield blk, cur
        cur = _resized(cur, resize_mode(blk.scale_factor))


def _predictor_plan(net, x, cond):
    backend = backend_default()
    batch, _, t = x.shape
    t_cond = cond.shape[-1] if cond is not None else None
    key = (batch, t, t_cond, x.device.index, backend)
    return net._plans.get(key, _signature(net), lambda: build_predictor_plan(net, batch, t, t_cond, backend))


def _check_input(x, channels: int):
    if x.dim() != 3 or x.shape[1] != channels:
        raise ValueError(f"expected an [N x {channels} x T] tensor, got {tuple(x.shape)}")


def check_predictor_inputs(net, x, cond, labels):
    """The argument contract of reference models/unet.py:126-131 plus the shape / range checks the reference gets for
    free from nn.Embedding and Conv1d (both raise on a bad label or channel count)."""
    assert (labels is None) == (net.num_labels is None), "must provide labels if and only if model is class conditional"
    assert (cond is None) == (net.cond_channels is None), "must provide cond sequence if and only if model is conditional"
    batch = x.shape[0]
    if labels is not None:
        if labels.dtype not in (torch.int64, torch.int32) or labels.numel() != batch:
            raise ValueError(f"labels must be {batch} integers, got {tuple(labels.shape)} {labels.dtype}")
        lo, hi = int(labels.min()), int(labels.max())
        if lo < 0 or hi >= net.num_labels:
            raise IndexError(f"label out of range: [{lo}, {hi}] with num_labels = {net.num_labels}")
    if cond is not None and (cond.dim() != 3 or cond.shape[0] != batch or cond.shape[1] != net.cond_channels):
        raise ValueError(f"cond must be [{batch} x {net.cond_channels} x T1], got {tuple(cond.shape)}")


def stage_predictor_inputs(net, plan: Plan, x, ts, cond, labels):
    """Point the program at this call's inputs (no copies for x; tiny copies for ts/labels/cond)."""
    check_predictor_inputs(net, x, cond, labels)
    plan.x_in = _f32(x)  # kept alive until the next call
    plan.slots["conv_in"].x = plan.x_in.data_ptr()
    plan.ts.copy_(ts.to(device=plan.device, dtype=torch.float32).reshape(-1).expand(plan.batch), non_blocking=True)
    if labels is not None:
        plan.labels.copy_(labels.to(plan.device).reshape(-1), non_blocking=True)
    if cond is not None:
        plan.cond_in.copy_(cond, non_blocking=True)


def predictor_forward(net, x, ts, cond=None, labels=None) -> torch.Tensor:
    _require_cuda(x, ts, cond, labels)
    _check_input(x, net.in_channels)
    with torch.no_grad():
        plan = _predictor_plan(net, x, cond)
        stage_predictor_inputs(net, plan, x, ts, cond, labels)
        if "conv_out" in plan.slots:
            plan.slots["conv_out"].mode = L.OUT_EPS
            plan.slots["conv_out"].out = plan.eps.data_ptr()
        plan.run()
        return plan.eps.clone()


# ---------------------------------------------------------------------------------------------
# UNetEncoder
# ---------------------------------------------------------------------------------------------
def build_encoder_plan(net, batch: int, t: int, backend: str) -> Plan:
    rate = net.downsample_rate
    if t % rate:
        raise ValueError(f"sequence length {t} must be divisible by the encoder downsample rate {rate}")
    if net.in_channels != 1:
        raise ValueError("the CUDA encoder path implements in_channels == 1 (waveforms)")
    device = next(net.parameters()).device
    _check_module(net)
    blocks = list(net.blocks)
    w = weights_for(net, blocks, backend)
    plan = Plan(device, batch, backend)
    plan.weights = w
    bc = net.base_channels
    alloc = _Alloc(plan, 2 * batch * (bc + 2 * sum(b.out_channels for b in blocks)))
    scratch = _scratch(plan, max(b.channels for b in blocks))
    max_elems = batch * bc * t
    for b in blocks:
        max_elems = max(max_elems, batch * b.out_channels * t)
    ping = [plan.empty(max_elems), plan.empty(max_elems)]
    h1_buf = plan.empty(max_elems)
    _memset_op(plan, alloc.arena)
    h = alloc.act(bc, t, ping[0])
    ci = L.ConvIn()
    ci.batch, ci.c_out, ci.t, ci.t_cond = batch, bc, t, 0
    ci.w, ci.bias = L.ptr(net.in_conv.weight), L.ptr(net.in_conv.bias)
    ci.out, ci.stats_out = h.ptr, h.stats_ptr
    plan.add(L.OP_CONV_IN, ci, "conv_in")
    flip, cur_t = 1, t
    for blk in blocks:
        t_out = _resized(cur_t, resize_mode(blk.scale_factor))
        h1 = alloc.act(blk.out_channels, t_out, h1_buf)
        out = alloc.act(blk.out_channels, t_out, ping[flip])
        _emit_block(plan, blk, [h], h1, out, w, scratch, None)
        h, cur_t, flip = out, t_out, flip ^ 1
    _emit_gn(plan, [h], net.out[0][0], scratch[0], scratch[1])
    plan.result = plan.empty(batch, net.out_channels, cur_t)
    _emit_conv(plan, [h], net.out[1], Act(plan.result, net.out_channels, cur_t, None), scale=scratch[0],
               shift=scratch[1], packed=w.packed.get((id(net), "out")))
    return plan.compile()


def encoder_forward(net, x) -> torch.Tensor:
    _require_cuda(x)
    _check_input(x, net.in_channels)
    backend = backend_default()
    batch, _, t = x.shape
    with torch.no_grad():
        plan = net._plans.get((batch, t, x.device.index, backend), _signature(net),
                              lambda: build_encoder_plan(net, batch, t, backend))
        plan.x_in = _f32(x)
        plan.slots["conv_in"].x = plan.x_in.data_ptr()
        plan.run()
        return plan.result.clone()


# ---------------------------------------------------------------------------------------------
# a single ResBlock (the reference exposes it as a module; also the unit of the parity tests)
# ---------------------------------------------------------------------------------------------
def run_single_block(blk, x, emb) -> torch.Tensor:
    _require_cuda(x, emb)
    _check_input(x, blk.channels)
    lib = L.load()
    backend = backend_default()
    batch, c, t = x.shape
    mode = resize_mode(blk.scale_factor)  # an odd t under pooling drops the last element, like avg_pool1d

    def build():
        w = weights_for(blk, [blk], backend)
        plan = Plan(x.device, batch, backend)
        plan.weights = w
        t_out = _resized(t, mode)
        alloc = _Alloc(plan, 2 * batch * (c + 2 * blk.out_channels))
        plan.src = alloc.act(c, t)
        h1 = alloc.act(blk.out_channels, t_out)
        plan.dst = alloc.act(blk.out_channels, t_out, stats=False)
        plan.arena = alloc.arena
        if blk.emb_channels:
            plan.gelu_emb = plan.empty(batch, blk.emb_channels)
            plan.ab = plan.empty(batch, w.film_total)
        _emit_block(plan, blk, [plan.src], h1, plan.dst, w, _scratch(plan, max(c, blk.out_channels)),
                    getattr(plan, "ab", None))
        return plan.compile()

    with torch.no_grad(), torch.cuda.device(x.device):
        plan = blk._plans.get((batch, t, x.device.index, backend, os.environ.get("VQVS_PREC")), _signature(blk), build)
        stream = L.stream_ptr(x.device)
        plan.arena.zero_()
        plan.src.buf.copy_(x)
        L.check(lib.vqvs_channel_stats(plan.src.ptr, batch, c, t, plan.src.stats_ptr, stream), "vqvs_channel_stats")
        if blk.emb_channels:
            e = _f32(emb)
            L.check(lib.vqvs_gelu(e.data_ptr(), plan.gelu_emb.data_ptr(), e.numel(), stream), "vqvs_gelu")
            w = plan.weights
            L.check(lib.vqvs_film_linear(plan.gelu_emb.data_ptr(), w.film_w.data_ptr(), w.film_b.data_ptr(), batch,
                                         blk.emb_channels, w.film_total, plan.ab.data_ptr(), stream), "vqvs_film_linear")
        plan.run()
        return plan.dst.buf.clone()


# ---------------------------------------------------------------------------------------------
# DDPM step (reference diffusion/diffusion.py:48-90) on top of the programs above
# ---------------------------------------------------------------------------------------------
def resolve_predictor(predictor):
    """Recognise our UNetPredictor behind the wrappers the reference's callers use
    (sample_diffusion.py:114 functools.partial(model.predictor, labels=...); BoundPredictor)."""
    import functools

    from .unet import UNetPredictor

    kwargs = {}
    fn = predictor
    while isinstance(fn, functools.partial):
        if fn.args:
            return None
        kwargs = {**fn.keywords, **kwargs}
        fn = fn.func
    if isinstance(fn, BoundPredictor):
        kwargs = {**fn.kwargs, **kwargs}
        fn = fn.net
    if not isinstance(fn, UNetPredictor) or fn.out_channels != 1:
        return None
    if set(kwargs) - {"cond", "labels", "use_checkpoint"}:
        return None
    return fn, kwargs.get("cond"), kwargs.get("labels")


class BoundPredictor:
    """predictor(xs, ts) with cond/labels bound -- what reference vq_vae.py:137-139 builds as a lambda,
    kept introspectable so the sampler can fuse the DDPM update into the network's last kernel."""

    def __init__(self, net, **kwargs):
        self.net, self.kwargs = net, kwargs

    def __call__(self, xs, ts, **extra):
        return self.net(xs, ts, **{**self.kwargs, **extra})


def ddpm_finish(x_t, eps, coef, noise, out, x0_sum=None):
    """out = c1*(x_t - c2*eps') + sigma*noise, eps' re-derived from the clamped x0 when x0_sum is given."""
    d = L.DdpmFinish()
    d.batch, d.t = x_t.shape[0], x_t[0].numel()
    d.use_x0_mean = 1 if x0_sum is not None else 0
    d.x_t, d.eps, d.noise, d.coef = x_t.data_ptr(), eps.data_ptr(), L.ptr(noise), coef.data_ptr()
    d.x0_sum, d.out = L.ptr(x0_sum), out.data_ptr()
    with torch.cuda.device(x_t.device):
        L.check(L.load().vqvs_ddpm_finish(C.byref(d), L.stream_ptr(x_t.device)), "vqvs_ddpm_finish")
    return out


def ddpm_x0_sum(x_t, eps, coef):
    s = torch.zeros(x_t.shape[0], dtype=torch.float64, device=x_t.device)
    with torch.cuda.device(x_t.device):
        L.check(L.load().vqvs_ddpm_x0_sum(x_t.data_ptr(), eps.data_ptr(), coef.data_ptr(), x_t.shape[0], x_t[0].numel(),
                                          s.data_ptr(), L.stream_ptr(x_t.device)), "vqvs_ddpm_x0_sum")
    return s


def fused_sample_step(net, x_t, ts, cond, labels, coef, noise, out, constrain: bool, want_eps: bool, first: bool):
    """One predictor forward whose final kernel applies the DDPM update.

    Returns `out` (x_{t-1}) for the plain and constrain cases, or the eps buffer when the caller
    has to run a cond_fn between eps and the update (want_eps)."""
    plan = _predictor_plan(net, x_t, cond)
    if first:
        stage_predictor_inputs(net, plan, x_t, ts, cond, labels)
    else:  # cond / labels are step-invariant: only x and ts move
        plan.x_in = x_t
        plan.slots["conv_in"].x = x_t.data_ptr()
        plan.ts.copy_(ts.reshape(-1).expand(plan.batch), non_blocking=True)
    co = plan.slots["conv_out"]
    co.x_t, co.coef = x_t.data_ptr(), coef.data_ptr()
    if want_eps:
        co.mode, co.out = L.OUT_EPS, plan.eps.data_ptr()
        plan.run()
        return plan.eps
    if constrain:
        plan.x0_sum.zero_()
        co.mode, co.out, co.x0_sum = L.OUT_X0_SUM, plan.eps.data_ptr(), plan.x0_sum.data_ptr()
        plan.run()
        return ddpm_finish(x_t, plan.eps, coef, noise, out, plan.x0_sum)
    co.mode, co.noise, co.out = L.OUT_PREV, L.ptr(noise), out.data_ptr()
    plan.run()
    return out
