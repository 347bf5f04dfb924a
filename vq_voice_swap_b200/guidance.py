"""Guidance gradients without ATen: forward and input-gradient programs of the reference's guidance models.

`cond_fn` of reference sample_diffusion.py:34-42 asks for d log p(label | x_t, t) / d x_t, and VQVAE.decode's
`enc_pred` guidance (vq_vae.py:125-130) for the gradient of EncoderPredictor.losses, both through `torch.autograd.grad`.
`ClassifierFunction` / `PredictorFunction` are torch.autograd.Functions whose forward runs the model as one libvqvs
program (every block's inputs x and intra-block tensor u stay resident) and whose backward runs a second program:

    for each ResBlock, last to first (reference models/unet.py:307-316 differentiated):
        dw  = conv2^T(dy)                      vqvs_conv1d_umma, weights transposed + taps flipped, same dilation
        du  = [GELU, FiLM, GN_b]^T(dw; u)      vqvs_gelu_bwd -> vqvs_gn_bwd_finalize -> vqvs_affine3
        dp  = conv1^T(du)                      vqvs_conv1d_umma
        dx  = [GELU, GN_a]^T(resize^T dp; x) + resize^T(skip^T(dy))      (skip^T = identity or the transposed 1x1 conv)
    A block fed by a channel concatenation (the UNet's up path) splits dp / the skip gradient by channel window; a tensor
    with two consumers (the UNet's skip stack) accumulates both gradients (VqvsAffine3.add2).

The classifier adds attention-pool^T and head^T in front and the input conv^T behind; the predictor adds the head conv^T
and (for EncoderPredictor) the strided sampling + 1x1 output conv.  No ATen kernel and no autograd graph is involved in
either direction; torch only owns the memory and the stream.
"""

import ctypes as C
from typing import List, Optional

import torch

from . import engine
from . import lib as L


def _pack_transposed(weight: torch.Tensor) -> Optional[engine.Packed]:
    """Operand image of the transposed conv: Wt[ci, co, k] = W[co, ci, K-1-k] (input gradient of a 'same' conv)."""
    wt = weight.detach().transpose(0, 1).flip(-1).contiguous()
    return engine.pack_weights(wt, None, L.PREC_BF16X3)


class Backward:
    """Emits the input-gradient program of a chain / DAG of ResBlocks into an engine.Plan."""

    def __init__(self, plan: engine.Plan, batch: int, saved, film_ab: Optional[torch.Tensor], film_offsets, film_total: int,
                 extra_elems: int = 0):
        self.plan, self.batch = plan, batch
        self.film_ab, self.film_offsets, self.film_total = film_ab, film_offsets, film_total
        c_max = 16
        elems = extra_elems
        n_gn = 2
        for _, srcs, u, out, _ in saved:
            c_in = sum(s.c for s in srcs)
            c_max = max(c_max, c_in, out.c)
            elems = max(elems, c_in * max(srcs[0].t, out.t), out.c * out.t)
            n_gn += 2
        self.c_max = c_max
        self.acc = plan.empty(n_gn * batch * c_max * 2, dtype=torch.float64)
        engine._memset_op(plan, self.acc)
        self.acc_used = 0
        self.prep = plan.empty(5 * batch * c_max)
        self.coef = plan.empty(3 * batch * c_max)
        self.scratch = [plan.empty(batch * elems) for _ in range(4)]  # dw/du, dp, skip^T(dy), q
        self.grads = {}     # id(Act) -> gradient tensor [batch, c, t]
        self.packed = []    # keep the transposed weight images alive

    # -- helpers ---------------------------------------------------------------------------------
    def grad_of(self, act) -> torch.Tensor:
        g = self.grads.get(id(act))
        if g is None:
            raise RuntimeError("backward reached a tensor that has no gradient yet (blocks must be visited last to first)")
        return g

    def _next_acc(self) -> torch.Tensor:
        n = self.batch * self.c_max * 2
        view = self.acc[self.acc_used:self.acc_used + n]
        self.acc_used += n
        assert self.acc_used <= self.acc.numel()
        return view

    def conv_t(self, src_ptr: int, c_in: int, c_out: int, t: int, ksize: int, dilation: int, weight: torch.Tensor, dst_ptr: int):
        packed = _pack_transposed(weight)
        if packed is None:
            raise ValueError(f"transposed conv {c_in}->{c_out}: channel counts must be multiples of 16")
        self.packed.append(packed)
        d = L.Conv()
        d.batch, d.c_a, d.c_b, d.t_in, d.c_out, d.t_out = self.batch, c_in, 0, t, c_out, t
        d.ksize, d.dilation, d.resize, d.act, d.skip_mode = ksize, dilation, L.RESIZE_NONE, 0, L.SKIP_NONE
        d.xa, d.out = src_ptr, dst_ptr
        d.w_packed = packed.img.data_ptr()
        d.reserved_ = packed.prec << L.CONV_PREC_SHIFT
        if not L.load().vqvs_conv1d_umma_supported(C.byref(d)):
            raise ValueError(f"transposed conv {c_in}->{c_out} k={ksize}: shape not supported by the tcgen05 kernel")
        self.plan.add(L.OP_CONV_UMMA, d)

    def gn_backward(self, fin, d_in_ptr: int, srcs: List, up: int, outs: List[torch.Tensor], add_ptr: int = 0, add_mode: int = 0):
        """[GELU, (FiLM), GroupNorm]^T over the channel concatenation of `srcs`.

        d_in [batch, sum c, t_d]: gradient w.r.t. resize(gelu(GN(cat(srcs)))); outs[k] receives d(srcs[k]) (and is ADDED to
        when that tensor already holds a gradient from another consumer); add: the skip path's gradient, same layout as d_in."""
        plan, batch = self.plan, self.batch
        c_total = sum(s.c for s in srcs)
        pp = L.GnBwdPrep()
        pp.gn, pp.prep = C.addressof(fin), self.prep.data_ptr()
        plan.add(L.OP_GN_BWD_PREP, pp)
        acc = self._next_acc()
        off = 0
        qs = []
        for k, s in enumerate(srcs):
            q = self.scratch[3] if len(srcs) == 1 else plan.empty(batch * s.c * s.t)
            qs.append(q)
            gb = L.GeluBwd()
            gb.batch, gb.c, gb.t, gb.up, gb.c_total, gb.c_off = batch, s.c, s.t, up, c_total, off
            gb.d_in, gb.z, gb.prep, gb.q, gb.acc = d_in_ptr, s.ptr, self.prep.data_ptr(), q.data_ptr(), acc.data_ptr()
            plan.add(L.OP_GELU_BWD, gb)
            off += s.c
        gf = L.GnBwdFinalize()
        gf.batch, gf.c, gf.groups, gf.count = batch, c_total, fin.groups, srcs[0].t
        gf.acc, gf.prep, gf.coef = acc.data_ptr(), self.prep.data_ptr(), self.coef.data_ptr()
        plan.add(L.OP_GN_BWD_FINALIZE, gf)
        off = 0
        for k, s in enumerate(srcs):
            af = L.Affine3()
            af.batch, af.c, af.t, af.add_mode, af.c_total, af.c_off = batch, s.c, s.t, add_mode, c_total, off
            af.q, af.z, af.coef, af.add = qs[k].data_ptr(), s.ptr, self.coef.data_ptr(), add_ptr
            prev = self.grads.get(id(s))
            af.add2 = prev.data_ptr() if prev is not None else 0
            af.out = outs[k].data_ptr()
            plan.add(L.OP_AFFINE3, af)
            self.grads[id(s)] = outs[k]
            off += s.c

    # -- one ResBlock ----------------------------------------------------------------------------
    def block(self, blk, srcs: List, u, out, mode: int):
        plan, batch = self.plan, self.batch
        dy = self.grad_of(out)
        bw, bp, bs = self.scratch[0], self.scratch[1], self.scratch[2]
        c_in = sum(s.c for s in srcs)
        conv1, conv2, proj = blk.pre_cond[2], engine._tail_conv(blk), engine._skip_proj(blk)
        # dw = conv2^T(dy); du = [GELU, FiLM, GN_b]^T(dw; u), in place
        self.conv_t(dy.data_ptr(), out.c, out.c, out.t, 3, conv2.dilation[0], conv2.weight, bw.data_ptr())
        film_ptr = film_stride = 0
        if getattr(blk, "emb_channels", None):
            film_ptr = self.film_ab.data_ptr() + 4 * self.film_offsets[id(blk)]
            film_stride = self.film_total
        fin_b = engine._emit_gn(plan, [u], blk.pre_cond[3], self.prep, self.prep, film_ptr, film_stride, standalone=False)
        du = bw[: batch * u.c * u.t].view(batch, u.c, u.t)
        saved_grad = self.grads.pop(id(u), None)  # u never has another consumer
        self.gn_backward(fin_b, bw.data_ptr(), [u], 0, [du])
        self.grads.pop(id(u), None)
        assert saved_grad is None
        # dp = conv1^T(du): gradient w.r.t. resize(gelu(GN_a(cat(srcs)))), length t_out
        self.conv_t(bw.data_ptr(), out.c, c_in, out.t, 3, 1, conv1.weight, bp.data_ptr())
        # skip path: ds = dy or skip^T(dy), length t_out, reaches the sources through resize^T
        if proj is not None:
            self.conv_t(dy.data_ptr(), out.c, c_in, out.t, 1, 1, proj.weight, bs.data_ptr())
            add = bs.data_ptr()
        else:
            add = dy.data_ptr()
        up = 1 if mode == L.RESIZE_DOWN2 else 2 if mode == L.RESIZE_UP2 else 0
        add_mode = 2 if mode == L.RESIZE_DOWN2 else 3 if mode == L.RESIZE_UP2 else 1
        fin_a = engine._emit_gn(plan, srcs, blk.pre_cond[0][0], self.prep, self.prep, standalone=False)
        outs = []
        for s in srcs:
            prev = self.grads.get(id(s))
            outs.append(prev if prev is not None else plan.empty(batch, s.c, s.t))
        self.gn_backward(fin_a, bp.data_ptr(), srcs, up, outs, add_ptr=add, add_mode=add_mode)


class _Plans:
    """Shared forward/backward bookkeeping: one forward, then (at most) its own backward."""

    generation = 0

    def _check_generation(self, generation: int):
        if generation != self.generation:
            raise RuntimeError("the model was evaluated again before this backward: its saved activations are gone "
                               "(one forward, then its backward -- the pattern of sample_diffusion.py's cond_fn)")


class GuidancePlans(_Plans):
    """Forward + backward launch programs of one Classifier (reference models/classifier.py) for one (batch, length)."""

    def __init__(self, clf, batch: int, t: int, backend: str):
        stem = clf.stem
        if t % (2 ** len(stem.channel_mult)):
            raise ValueError(f"sequence length {t} must be divisible by {2 ** len(stem.channel_mult)} (one halving per level)")
        engine._check_module(clf)
        self.clf, self.batch, self.t = clf, batch, t
        device = next(clf.parameters()).device
        self.device = device
        blocks = list(stem.blocks)
        w = engine.weights_for(stem, blocks, backend)
        fwd = engine.Plan(device, batch, backend)
        fwd.weights = w
        bc = stem.base_channels
        emb_dim = stem.embed_dim
        alloc = engine._Alloc(fwd, 2 * batch * (bc + 2 * sum(b.out_channels for b in blocks)))
        scratch = engine._scratch(fwd, max(max(b.channels, b.out_channels) for b in blocks))
        fwd.ts = fwd.empty(batch)
        fwd.emb = fwd.empty(batch, emb_dim)
        fwd.gelu_emb = fwd.empty(batch, emb_dim)
        fwd.ab = fwd.empty(batch, w.film_total)

        engine._memset_op(fwd, alloc.arena)
        te = L.TimeEmbed()
        te.batch, te.dim = batch, emb_dim
        te.ts, te.freqs = fwd.ts.data_ptr(), w.freqs.data_ptr()
        te.w1, te.b1 = L.ptr(stem.time_embed.proj.weight), L.ptr(stem.time_embed.proj.bias)
        te.w2, te.b2 = L.ptr(stem.time_embed_extra[1].weight), L.ptr(stem.time_embed_extra[1].bias)
        te.emb, te.gelu_emb = fwd.emb.data_ptr(), fwd.gelu_emb.data_ptr()
        fwd.add(L.OP_TIME_EMBED, te)
        fl = L.Film()
        fl.gelu_emb, fl.w_cat, fl.b_cat = fwd.gelu_emb.data_ptr(), w.film_w.data_ptr(), w.film_b.data_ptr()
        fl.batch, fl.dim, fl.n_out, fl.ab = batch, emb_dim, w.film_total, fwd.ab.data_ptr()
        fwd.add(L.OP_FILM, fl)

        h = alloc.act(bc, t)
        self.h0 = h
        ci = L.ConvIn()
        ci.batch, ci.c_out, ci.t, ci.t_cond = batch, bc, t, 0
        ci.w, ci.bias = L.ptr(stem.in_conv.weight), L.ptr(stem.in_conv.bias)
        ci.out, ci.stats_out = h.ptr, h.stats_ptr
        fwd.add(L.OP_CONV_IN, ci, "conv_in")

        # every block keeps its input x and its intra-block tensor u: the backward program reads both
        self.saved = []
        cur_t = t
        for blk in blocks:
            mode = engine.resize_mode(blk.scale_factor)
            t_out = engine._resized(cur_t, mode)
            u = alloc.act(blk.out_channels, t_out)
            out = alloc.act(blk.out_channels, t_out)
            engine._emit_block(fwd, blk, [h], u, out, w, scratch, fwd.ab)
            self.saved.append((blk, [h], u, out, mode))
            h, cur_t = out, t_out
        self.h_last, self.t_last = h, cur_t

        # GroupNorm + GELU + attention pool + head
        pool = stem.out[1]
        c_last = h.c
        self.prep_final = fwd.empty(5 * batch * c_last)
        self.fin_final = engine._emit_gn(fwd, [h], stem.out[0][0], scratch[0], scratch[1], standalone=False)
        pp = L.GnBwdPrep()
        pp.gn, pp.prep = C.addressof(self.fin_final), self.prep_final.data_ptr()
        fwd.add(L.OP_GN_BWD_PREP, pp)
        ws_bytes = L.load().vqvs_attnpool_workspace_bytes(batch, c_last, cur_t, pool.num_heads)
        if ws_bytes <= 0:
            raise ValueError("attention pool: unsupported shape")
        self.ap_ws = fwd.empty(ws_bytes // 4)
        self.stem_out = fwd.empty(batch, stem.out_channels)
        fwd.add(L.OP_ATTNPOOL_FWD, self._attnpool())
        self.logits = fwd.empty(batch, clf.num_labels)
        fwd.add(L.OP_CLS_HEAD_FWD, self._head())
        self.fwd = fwd.compile()
        self._build_backward(w, backend)

    def _attnpool(self) -> L.AttnPool:
        pool, stem = self.clf.stem.out[1], self.clf.stem
        ap = L.AttnPool()
        ap.batch, ap.c, ap.t, ap.heads, ap.c_out = self.batch, self.h_last.c, self.t_last, pool.num_heads, stem.out_channels
        ap.h, ap.prep = self.h_last.ptr, self.prep_final.data_ptr()
        ap.w_qkv, ap.b_qkv = L.ptr(pool.qkv_proj.weight), L.ptr(pool.qkv_proj.bias)
        ap.w_proj, ap.b_proj = L.ptr(pool.c_proj.weight), L.ptr(pool.c_proj.bias)
        ap.ws, ap.out = self.ap_ws.data_ptr(), self.stem_out.data_ptr()
        return ap

    def _head(self) -> L.ClsHead:
        head = self.clf.out[1]
        hd = L.ClsHead()
        hd.batch, hd.dim, hd.labels = self.batch, self.clf.stem.out_channels, self.clf.num_labels
        hd.stem, hd.w, hd.b = self.stem_out.data_ptr(), L.ptr(head.weight), L.ptr(head.bias)
        hd.logits = self.logits.data_ptr()
        return hd

    def _build_backward(self, w, backend):
        clf, stem, batch = self.clf, self.clf.stem, self.batch
        bwd = engine.Plan(self.device, batch, backend)
        c_last, t_last = self.h_last.c, self.t_last
        bk = Backward(bwd, batch, self.saved, self.fwd.ab, w.film_offsets, w.film_total, extra_elems=c_last * t_last)
        self.d_logits = bwd.empty(batch, clf.num_labels)
        d_stem = bwd.empty(batch, stem.out_channels)
        self.dx = bwd.empty(batch, 1, self.t)
        hd = self._head()
        hd.d_logits, hd.d_stem = self.d_logits.data_ptr(), d_stem.data_ptr()
        bwd.add(L.OP_CLS_HEAD_BWD, hd)
        d_act = bwd.empty(batch, c_last, t_last)
        ap = self._attnpool()
        ap.d_out, ap.d_act = d_stem.data_ptr(), d_act.data_ptr()
        bwd.add(L.OP_ATTNPOOL_BWD, ap)
        bk.gn_backward(self.fin_final, d_act.data_ptr(), [self.h_last], 0, [bwd.empty(batch, c_last, t_last)])
        for blk, srcs, u, out, mode in reversed(self.saved):
            bk.block(blk, srcs, u, out, mode)
        cb = L.ConvInBwd()
        cb.batch, cb.c, cb.t = batch, stem.base_channels, self.t
        cb.dh, cb.w, cb.dx = bk.grad_of(self.h0).data_ptr(), L.ptr(stem.in_conv.weight), self.dx.data_ptr()
        bwd.add(L.OP_CONV_IN_BWD, cb)
        self.bk = bk
        self.bwd = bwd.compile()

    # -----------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
        self.x_in = engine._f32(x)
        self.fwd.slots["conv_in"].x = self.x_in.data_ptr()
        self.fwd.ts.copy_(ts.to(device=self.device, dtype=torch.float32).reshape(-1).expand(self.batch), non_blocking=True)
        self.fwd.run()
        self.generation += 1
        return self.logits.clone()

    def backward(self, d_logits: torch.Tensor, generation: int) -> torch.Tensor:
        self._check_generation(generation)
        self.d_logits.copy_(d_logits.to(self.d_logits), non_blocking=True)
        self.bwd.run()
        return self.dx.clone()


class PredictorGuidancePlans(_Plans):
    """Forward + backward programs of a UNetPredictor with `out_channels` feature maps, optionally followed by the strided
    sampling and 1x1 conv of EncoderPredictor (reference models/encoder_predictor.py:40-59)."""

    def __init__(self, net, batch: int, t: int, backend: str, rate: int = 1, out_conv: Optional[torch.nn.Conv1d] = None):
        if net.cond_channels is not None or net.num_labels is not None:
            raise NotImplementedError("guidance through a conditional UNetPredictor is not implemented")
        if net.out_channels % 16:
            raise ValueError("the predictor's out_channels must be a multiple of 16 for the tcgen05 head conv")
        self.net, self.batch, self.t, self.rate, self.out_conv = net, batch, t, rate, out_conv
        self.device = next(net.parameters()).device
        fwd = engine.build_predictor_plan(net, batch, t, None, backend, keep_activations=True)
        self.fwd = fwd
        lib = L.load()
        head_gn, head_conv = net.out[0][0], net.out[1]
        bwd = engine.Plan(self.device, batch, backend)
        w = fwd.weights
        bk = Backward(bwd, batch, fwd.saved, fwd.ab, w.film_offsets, w.film_total, extra_elems=max(net.out_channels, net.base_channels) * t)
        self.features = fwd.eps                      # [batch, out_channels, t]
        self.d_features = bwd.empty(batch, net.out_channels, t)
        self.dx = bwd.empty(batch, 1, t)
        if out_conv is not None:                     # logits = out_conv(features[:, :, ::rate])
            t1 = t // rate
            self.sampled = torch.empty(batch, net.out_channels, t1, device=self.device)
            self.logits = torch.empty(batch, out_conv.out_channels, t1, device=self.device)
            self.d_logits = bwd.empty(batch, out_conv.out_channels, t1)
            self.d_sampled = bwd.empty(batch, net.out_channels, t1)
            self.packed_out = engine.pack_weights(out_conv.weight, None, L.PREC_BF16X3)
            if self.packed_out is None:
                raise ValueError("EncoderPredictor: num_latents and bottleneck_dim must be multiples of 16")
            bk.conv_t(self.d_logits.data_ptr(), out_conv.out_channels, net.out_channels, t1, 1, 1, out_conv.weight, self.d_sampled.data_ptr())
            self._scatter_in_backward = True
        else:
            self.d_logits = self.d_features
            self._scatter_in_backward = False
        # head conv^T: d(gelu(GN(h_last))) = conv^T(d_features); then the GroupNorm in front of it
        d_act = bwd.empty(batch, net.base_channels, t)
        self._head_conv_t_at = len(bwd.descs)
        bk.conv_t(self.d_features.data_ptr(), net.out_channels, net.base_channels, t, 3, 1, head_conv.weight, d_act.data_ptr())
        fin_head = engine._emit_gn(bwd, [fwd.h_last], head_gn, bk.prep, bk.prep, standalone=False)
        bk.gn_backward(fin_head, d_act.data_ptr(), [fwd.h_last], 0, [bwd.empty(batch, fwd.h_last.c, fwd.h_last.t)])
        for blk, srcs, u, out, mode in reversed(fwd.saved):
            bk.block(blk, srcs, u, out, mode)
        cb = L.ConvInBwd()
        cb.batch, cb.c, cb.t = batch, net.base_channels, t
        cb.dh, cb.w, cb.dx = bk.grad_of(fwd.h0).data_ptr(), L.ptr(net.in_conv.weight), self.dx.data_ptr()
        bwd.add(L.OP_CONV_IN_BWD, cb)
        self.bk = bk
        self.bwd = bwd.compile()
        self._lib = lib

    def forward(self, x: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
        plan, net = self.fwd, self.net
        engine.stage_predictor_inputs(net, plan, x, ts, None, None)
        plan.run()
        self.generation += 1
        if self.out_conv is None:
            return self.features.clone()
        stream = L.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            L.check(self._lib.vqvs_stride_sample(self.features.data_ptr(), self.sampled.data_ptr(), self.batch * net.out_channels,
                                                 self.t, self.rate, 0, stream), "vqvs_stride_sample")
            d = L.Conv()
            d.batch, d.c_a, d.t_in, d.c_out, d.t_out = self.batch, net.out_channels, self.t // self.rate, self.out_conv.out_channels, self.t // self.rate
            d.ksize, d.dilation = 1, 1
            d.xa, d.out, d.bias = self.sampled.data_ptr(), self.logits.data_ptr(), L.ptr(self.out_conv.bias)
            d.w_packed = self.packed_out.img.data_ptr()
            d.reserved_ = self.packed_out.prec << L.CONV_PREC_SHIFT
            L.check(self._lib.vqvs_conv1d_umma(C.byref(d), stream), "vqvs_conv1d_umma")
        return self.logits.clone()

    def backward(self, d_out: torch.Tensor, generation: int) -> torch.Tensor:
        self._check_generation(generation)
        self.d_logits.copy_(d_out.to(self.d_logits), non_blocking=True)
        if self._scatter_in_backward:
            # the program's first two ops are the acc memset and out_conv^T; the strided scatter sits between out_conv^T and
            # the head conv^T, so the program runs in two pieces around it
            n0 = self._head_conv_t_at
            stream = L.stream_ptr(self.device)
            with torch.cuda.device(self.device):
                L.check(self._lib.vqvs_run(self.bwd.ops, n0, stream), "vqvs_run")
                L.check(self._lib.vqvs_stride_sample(self.d_sampled.data_ptr(), self.d_features.data_ptr(),
                                                     self.batch * self.net.out_channels, self.t, self.rate, 1, stream), "vqvs_stride_sample")
                rest = C.cast(C.byref(self.bwd.ops, n0 * C.sizeof(L.Op)), C.POINTER(L.Op))
                L.check(self._lib.vqvs_run(rest, len(self.bwd.descs) - n0, stream), "vqvs_run")
        else:
            self.bwd.run()
        return self.dx.clone()


def classifier_plans(clf, x: torch.Tensor) -> GuidancePlans:
    engine._require_cuda(x)
    engine._check_input(x, 1)
    backend = engine.backend_default()
    batch, _, t = x.shape
    key = (batch, t, x.device.index, backend)
    return clf._plans.get(key, engine._signature(clf), lambda: GuidancePlans(clf, batch, t, backend))


class ClassifierFunction(torch.autograd.Function):
    """logits = Classifier(x, ts); backward returns d(loss)/dx from the native dgrad program (ts gets no gradient)."""

    @staticmethod
    def forward(ctx, x, ts, clf):
        plans = classifier_plans(clf, x)
        logits = plans.forward(x, ts)
        ctx.plans, ctx.generation, ctx.x_dtype = plans, plans.generation, x.dtype
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        with torch.no_grad():
            dx = ctx.plans.backward(d_logits.contiguous(), ctx.generation)
        return dx.to(ctx.x_dtype), None, None


class PredictorFunction(torch.autograd.Function):
    """EncoderPredictor.forward: logits [N x D x T/R] with a native backward to x."""

    @staticmethod
    def forward(ctx, x, ts, owner):
        engine._require_cuda(x, ts)
        engine._check_input(x, 1)
        backend = engine.backend_default()
        batch, _, t = x.shape
        key = (batch, t, x.device.index, backend)
        plans = owner._plans.get(key, engine._signature(owner), lambda: PredictorGuidancePlans(
            owner.unet, batch, t, backend, rate=owner.downsample_rate, out_conv=owner.out))
        out = plans.forward(x, ts)
        ctx.plans, ctx.generation, ctx.x_dtype = plans, plans.generation, x.dtype
        return out

    @staticmethod
    def backward(ctx, d_out):
        with torch.no_grad():
            dx = ctx.plans.backward(d_out.contiguous(), ctx.generation)
        return dx.to(ctx.x_dtype), None, None
