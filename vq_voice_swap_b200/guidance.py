"""Classifier guidance without ATen: forward and input-gradient programs of reference models/classifier.py.

`cond_fn` of reference sample_diffusion.py:34-42 asks for d log p(label | x_t, t) / d x_t through `torch.autograd.grad`.
`ClassifierFunction` is a torch.autograd.Function whose forward runs the guidance model as one libvqvs program
(time embedding, FiLM table, input conv, 27 FiLM ResBlocks on the tcgen05 conv kernel, GroupNorm/GELU, attention pool,
Linear head) keeping every block's input x and intra-block tensor u resident, and whose backward runs a second program:

    d_logits -> head^T -> attention-pool^T -> [GELU, GroupNorm]^T -> for each ResBlock, last to first:
        dw  = conv2^T(dy)                      vqvs_conv1d_umma, weights transposed + taps flipped, same dilation
        du  = [GELU, FiLM, GN_b]^T(dw; u)      vqvs_gelu_bwd -> vqvs_gn_bwd_finalize -> vqvs_affine3
        dp  = conv1^T(du)                      vqvs_conv1d_umma
        dx  = [GELU, GN_a]^T(resize^T dp; x) + resize^T(skip^T(dy))      (skip^T = identity or the transposed 1x1 conv)
    -> input conv^T -> dx [N, 1, T]

No ATen kernel and no autograd graph is involved in either direction; torch only owns the memory and the stream.
"""

import ctypes as C
import math
from typing import List, Optional

import torch

from . import engine
from . import lib as L


def _pack_transposed(weight: torch.Tensor) -> Optional[engine.Packed]:
    """Operand image of the transposed conv: Wt[ci, co, k] = W[co, ci, K-1-k] (input gradient of a 'same' conv)."""
    wt = weight.detach().transpose(0, 1).flip(-1).contiguous()
    return engine.pack_weights(wt, None, L.PREC_BF16X3)


class GuidancePlans:
    """Forward + backward launch programs of one Classifier for one (batch, length)."""

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
        ci = L.ConvIn()
        ci.batch, ci.c_out, ci.t, ci.t_cond = batch, bc, t, 0
        ci.w, ci.bias = L.ptr(stem.in_conv.weight), L.ptr(stem.in_conv.bias)
        ci.out, ci.stats_out = h.ptr, h.stats_ptr
        fwd.add(L.OP_CONV_IN, ci, "conv_in")

        # every block keeps its input x and its intra-block tensor u: the backward program reads both
        self.saved = []  # (block, x Act, u Act, out Act, resize mode)
        cur_t = t
        for blk in blocks:
            mode = engine.resize_mode(blk.scale_factor)
            t_out = engine._resized(cur_t, mode)
            u = alloc.act(blk.out_channels, t_out)
            out = alloc.act(blk.out_channels, t_out)
            engine._emit_block(fwd, blk, [h], u, out, w, scratch, fwd.ab)
            self.saved.append((blk, h, u, out, mode))
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
        heads = pool.num_heads
        ws_bytes = L.load().vqvs_attnpool_workspace_bytes(batch, c_last, cur_t, heads)
        if ws_bytes <= 0:
            raise ValueError("attention pool: unsupported shape")
        self.ap_ws = fwd.empty(ws_bytes // 4)
        self.stem_out = fwd.empty(batch, stem.out_channels)
        ap = L.AttnPool()
        ap.batch, ap.c, ap.t, ap.heads, ap.c_out = batch, c_last, cur_t, heads, stem.out_channels
        ap.h, ap.prep = h.ptr, self.prep_final.data_ptr()
        ap.w_qkv, ap.b_qkv = L.ptr(pool.qkv_proj.weight), L.ptr(pool.qkv_proj.bias)
        ap.w_proj, ap.b_proj = L.ptr(pool.c_proj.weight), L.ptr(pool.c_proj.bias)
        ap.ws, ap.out = self.ap_ws.data_ptr(), self.stem_out.data_ptr()
        fwd.add(L.OP_ATTNPOOL_FWD, ap)
        head = clf.out[1]
        self.logits = fwd.empty(batch, clf.num_labels)
        hd = L.ClsHead()
        hd.batch, hd.dim, hd.labels = batch, stem.out_channels, clf.num_labels
        hd.stem, hd.w, hd.b = self.stem_out.data_ptr(), L.ptr(head.weight), L.ptr(head.bias)
        hd.logits = self.logits.data_ptr()
        fwd.add(L.OP_CLS_HEAD_FWD, hd)
        self.fwd = fwd.compile()
        self.generation = 0
        self._build_backward(w, backend)

    # -----------------------------------------------------------------------------------------
    def _build_backward(self, w, backend):
        clf, stem, batch = self.clf, self.clf.stem, self.batch
        bwd = engine.Plan(self.device, batch, backend)
        saved = self.saved
        c_max = max(max(x.c, out.c) for _, x, _, out, _ in saved)
        n_gn = 2 * len(saved) + 1
        acc = bwd.empty(n_gn * batch * c_max * 2, dtype=torch.float64)
        engine._memset_op(bwd, acc)
        prep = bwd.empty(5 * batch * c_max)
        coef = bwd.empty(3 * batch * c_max)
        max_elems = batch * max(max(x.c * x.t, out.c * out.t) for _, x, _, out, _ in saved)
        bufs = [bwd.empty(max_elems) for _ in range(4)]
        self.d_logits = bwd.empty(batch, clf.num_labels)
        d_stem = bwd.empty(batch, stem.out_channels)
        self.dx = bwd.empty(batch, 1, self.t)
        acc_slot = [0]

        def next_acc(c):
            view = acc[acc_slot[0]:acc_slot[0] + batch * c * 2]
            acc_slot[0] += batch * c_max * 2
            return view

        def gn_backward(fin, d_in, z, c, t, up, q_out, out, add=None, add_mode=0):
            """[GELU, (FiLM), GroupNorm]^T: out = d(z) given d_in = gradient w.r.t. gelu(GN(z))."""
            pp = L.GnBwdPrep()
            pp.gn, pp.prep = C.addressof(fin), prep.data_ptr()
            bwd.add(L.OP_GN_BWD_PREP, pp)
            a = next_acc(c)
            gb = L.GeluBwd()
            gb.batch, gb.c, gb.t, gb.up = batch, c, t, up
            gb.d_in, gb.z, gb.prep, gb.q, gb.acc = d_in, z, prep.data_ptr(), q_out, a.data_ptr()
            bwd.add(L.OP_GELU_BWD, gb)
            gf = L.GnBwdFinalize()
            gf.batch, gf.c, gf.groups, gf.count = batch, c, fin.groups, t
            gf.acc, gf.prep, gf.coef = a.data_ptr(), prep.data_ptr(), coef.data_ptr()
            bwd.add(L.OP_GN_BWD_FINALIZE, gf)
            af = L.Affine3()
            af.batch, af.c, af.t, af.add_mode = batch, c, t, add_mode
            af.q, af.z, af.coef, af.add, af.out = q_out, z, coef.data_ptr(), add or 0, out
            bwd.add(L.OP_AFFINE3, af)

        def conv_t(src_ptr, c_in, c_out, t, ksize, dilation, packed, dst_ptr):
            d = L.Conv()
            d.batch, d.c_a, d.c_b, d.t_in, d.c_out, d.t_out = batch, c_in, 0, t, c_out, t
            d.ksize, d.dilation, d.resize, d.act, d.skip_mode = ksize, dilation, L.RESIZE_NONE, 0, L.SKIP_NONE
            d.xa, d.out = src_ptr, dst_ptr
            d.w_packed = packed.img.data_ptr()
            d.reserved_ = packed.prec << L.CONV_PREC_SHIFT
            if not L.load().vqvs_conv1d_umma_supported(C.byref(d)):
                raise ValueError(f"transposed conv {c_in}->{c_out} k={ksize}: shape not supported by the tcgen05 kernel")
            bwd.add(L.OP_CONV_UMMA, d)

        # head^T, attention pool^T
        head = clf.out[1]
        hd = L.ClsHead()
        hd.batch, hd.dim, hd.labels = batch, stem.out_channels, clf.num_labels
        hd.stem, hd.w, hd.b = self.stem_out.data_ptr(), L.ptr(head.weight), L.ptr(head.bias)
        hd.d_logits, hd.d_stem = self.d_logits.data_ptr(), d_stem.data_ptr()
        bwd.add(L.OP_CLS_HEAD_BWD, hd)
        pool = stem.out[1]
        c_last, t_last = self.h_last.c, self.t_last
        ap = L.AttnPool()
        ap.batch, ap.c, ap.t, ap.heads, ap.c_out = batch, c_last, t_last, pool.num_heads, stem.out_channels
        ap.h, ap.prep = self.h_last.ptr, self.prep_final.data_ptr()
        ap.w_qkv, ap.b_qkv = L.ptr(pool.qkv_proj.weight), L.ptr(pool.qkv_proj.bias)
        ap.w_proj, ap.b_proj = L.ptr(pool.c_proj.weight), L.ptr(pool.c_proj.bias)
        ap.ws, ap.d_out, ap.d_act = self.ap_ws.data_ptr(), d_stem.data_ptr(), bufs[1].data_ptr()
        bwd.add(L.OP_ATTNPOOL_BWD, ap)
        # final GroupNorm + GELU: gradient w.r.t. the last block's output lands in bufs[0]
        gn_backward(self.fin_final, bufs[1].data_ptr(), self.h_last.ptr, c_last, t_last, 0, bufs[1].data_ptr(), bufs[0].data_ptr())
        dy = 0  # index of the buffer holding the current gradient

        self.packed_t = []
        for blk, x, u, out, mode in reversed(saved):
            free = [i for i in range(4) if i != dy]
            b_w, b_p, b_q = free
            pool_blk = mode == L.RESIZE_DOWN2
            conv1, conv2, proj = blk.pre_cond[2], engine._tail_conv(blk), engine._skip_proj(blk)
            pk2, pk1 = _pack_transposed(conv2.weight), _pack_transposed(conv1.weight)
            pks = _pack_transposed(proj.weight) if proj is not None else None
            self.packed_t += [pk2, pk1, pks]
            # dw = conv2^T(dy); du = [GELU, FiLM, GN_b]^T(dw; u), in place
            conv_t(bufs[dy].data_ptr(), out.c, out.c, out.t, 3, conv2.dilation[0], pk2, bufs[b_w].data_ptr())
            film_ptr = self.fwd.ab.data_ptr() + 4 * self.fwd.weights.film_offsets[id(blk)]
            fin_b = engine._emit_gn(bwd, [u], blk.pre_cond[3], prep, prep, film_ptr, self.fwd.weights.film_total, standalone=False)
            gn_backward(fin_b, bufs[b_w].data_ptr(), u.ptr, u.c, u.t, 0, bufs[b_w].data_ptr(), bufs[b_w].data_ptr())
            # dp = conv1^T(du)
            conv_t(bufs[b_w].data_ptr(), out.c, x.c, out.t, 3, 1, pk1, bufs[b_p].data_ptr())
            # skip path: ds = dy or skip^T(dy) (length t_out), reaches x through resize^T
            if pks is not None:
                conv_t(bufs[dy].data_ptr(), out.c, x.c, out.t, 1, 1, pks, bufs[b_w].data_ptr())
                add = bufs[b_w].data_ptr()
            else:
                add = bufs[dy].data_ptr()
            # dx = [GELU, GN_a]^T(resize^T dp; x) + resize^T ds
            fin_a = engine._emit_gn(bwd, [x], blk.pre_cond[0][0], prep, prep, standalone=False)
            gn_backward(fin_a, bufs[b_p].data_ptr(), x.ptr, x.c, x.t, 1 if pool_blk else 0, bufs[b_q].data_ptr(), bufs[b_q].data_ptr(),
                        add=add, add_mode=2 if pool_blk else 1)
            dy = b_q
        cb = L.ConvInBwd()
        cb.batch, cb.c, cb.t = batch, stem.base_channels, self.t
        cb.dh, cb.w, cb.dx = bufs[dy].data_ptr(), L.ptr(stem.in_conv.weight), self.dx.data_ptr()
        bwd.add(L.OP_CONV_IN_BWD, cb)
        self.bwd = bwd.compile()

    # -----------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
        self.x_in = engine._f32(x)
        self.fwd.slots["conv_in"].x = self.x_in.data_ptr()
        self.fwd.ts.copy_(ts.to(device=self.device, dtype=torch.float32).reshape(-1).expand(self.batch), non_blocking=True)
        self.fwd.run()
        self.generation += 1
        return self.logits.clone()

    def stem_features(self) -> torch.Tensor:
        return self.stem_out.clone()

    def backward(self, d_logits: torch.Tensor, generation: int) -> torch.Tensor:
        if generation != self.generation:
            raise RuntimeError("the classifier was evaluated again before this backward: its saved activations are gone "
                               "(one forward, then its backward -- the pattern of sample_diffusion.py's cond_fn)")
        self.d_logits.copy_(d_logits.to(self.d_logits), non_blocking=True)
        self.bwd.run()
        return self.dx.clone()


def plans_for(clf, x: torch.Tensor) -> GuidancePlans:
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
        plans = plans_for(clf, x)
        logits = plans.forward(x, ts)
        ctx.plans, ctx.generation, ctx.x_dtype = plans, plans.generation, x.dtype
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        with torch.no_grad():
            dx = ctx.plans.backward(d_logits.contiguous(), ctx.generation)
        return dx.to(ctx.x_dtype), None, None


def attention_heads(channels: int, head_channels: int) -> int:
    return channels // head_channels


def _unused(*a):  # keep linters quiet about typing-only imports
    return List, math
