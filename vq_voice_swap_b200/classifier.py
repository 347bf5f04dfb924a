"""Noised-audio classifier used for classifier guidance (reference models/classifier.py).

`Classifier` / `ClassifierStem` / `AttentionPool1d` own parameters under the reference's
state-dict names (`stem.blocks.{i}.pre_cond.2.weight`, `stem.out.1.qkv_proj.weight`, ...) and are
created in the reference's order, so a torch seed initialises both identically and reference
checkpoints load.

The only caller on the sampling path is `cond_fn` (reference sample_diffusion.py:34-42), which needs
d log p(label | x_t, t) / d x_t through `torch.autograd.grad`.  `Classifier.forward` is a
torch.autograd.Function (guidance.ClassifierFunction) whose forward AND backward are libvqvs launch
programs: the stem's convolutions -- forward and transposed -- on the tcgen05 kernel, GroupNorm/GELU/
FiLM backward, attention pool and head in hand-written CUDA (csrc/guidance.cu).  No ATen kernel runs in
either direction.  `forward_aten` below restates the module with torch ops; it is a diagnostic (parameter
layout cross-check on any device, A/B timing with VQVS_GUIDANCE=aten) that no product path selects.
"""

import os

import math
from typing import Any, Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine
from .base import Savable
from .unet import DEFAULT_MULT, ResBlock, TimeEmbedding, UNetPredictor, _Slot, _scaled_, group_count

GUIDANCE_ENGINE = "libvqvs forward + dgrad programs (guidance.ClassifierFunction)"


def _use_aten() -> bool:
    """VQVS_GUIDANCE=aten: evaluate the guidance model with ATen under autograd (A/B measurement only)."""
    return os.environ.get("VQVS_GUIDANCE", "native") == "aten"


# ---------------------------------------------------------------------------------------------
# autograd evaluation of the shared building blocks (same parameters, ATen ops)
# ---------------------------------------------------------------------------------------------
def _resize(x: torch.Tensor, factor: float) -> torch.Tensor:
    """reference models/unet.py:324-334."""
    if factor == 1.0:
        return x
    if factor == 0.5:
        return F.avg_pool1d(x, 2)
    if factor == 2.0:
        return x.repeat_interleave(2, dim=-1)
    raise ValueError(f"unsupported scale factor {factor}")


def _group_norm(x: torch.Tensor, gn: nn.GroupNorm) -> torch.Tensor:
    return F.group_norm(x, gn.num_groups, gn.weight, gn.bias, gn.eps)


def _conv(x: torch.Tensor, conv: nn.Conv1d) -> torch.Tensor:
    return F.conv1d(x, conv.weight, conv.bias, padding=conv.padding[0], dilation=conv.dilation[0])


def resblock_autograd(blk: ResBlock, x: torch.Tensor, emb: Optional[torch.Tensor]) -> torch.Tensor:
    """reference models/unet.py:307-316 on the block's own parameters, differentiable."""
    h = F.gelu(_group_norm(x, blk.pre_cond[0][0]))
    h = _conv(_resize(h, blk.scale_factor), blk.pre_cond[2])
    h = _group_norm(h, blk.pre_cond[3])
    if blk.emb_channels:
        film = blk.cond_layers[1]
        ab = F.linear(F.gelu(emb), film.weight, film.bias)
        gain, offset = ab[:, : blk.out_channels, None], ab[:, blk.out_channels:, None]
        h = h * (gain + 1) + offset
    h = F.gelu(h)
    if blk.dropout and blk.training:
        h = F.dropout(h, blk.dropout)
    h = _conv(h, blk.post_cond[len(blk.post_cond) - 1])
    s = _resize(x, blk.scale_factor)
    proj = blk.skip[1]
    if isinstance(proj, nn.Conv1d):
        s = _conv(s, proj)
    return s + h


def time_embedding_autograd(te: TimeEmbedding, extra: nn.Sequential, ts: torch.Tensor) -> torch.Tensor:
    """reference models/wavegrad.py:359-373 followed by models/unet.py:41-43."""
    half = te.channels // 2
    freqs = (
        torch.exp(-math.log(100.0 / 0.1) * torch.arange(start=0, end=half, dtype=torch.float32) / (half - 1)) * 100.0
    ).to(ts)
    args = ts[:, None] * freqs[None]
    e = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    e = F.linear(e, te.proj.weight, te.proj.bias)
    lin = extra[1]
    return F.linear(F.gelu(e), lin.weight, lin.bias)


# ---------------------------------------------------------------------------------------------
# modules
# ---------------------------------------------------------------------------------------------
class AttentionPool1d(nn.Module):
    """reference models/classifier.py:133-158: prepend a zero token, 1x1 qkv projection, multi-head
    attention, 1x1 output projection, keep position 0.  Only the first query row is ever consumed,
    so only that row of the attention matrix is evaluated."""

    def __init__(self, channels: int, head_channels: int = 64, out_channels: Optional[int] = None):
        super().__init__()
        assert channels % head_channels == 0, f"head channels ({head_channels}) must divide channels ({channels})"
        self.qkv_proj = nn.Conv1d(channels, 3 * channels, 1)
        self.c_proj = nn.Conv1d(channels, out_channels or channels, 1)
        self.num_heads = channels // head_channels

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, c, _ = x.shape
        heads, ch = self.num_heads, c // self.num_heads
        tokens = F.pad(x, (1, 0))                                            # zero token first: [n, c, t+1]
        qkv = F.conv1d(tokens, self.qkv_proj.weight, self.qkv_proj.bias)     # [n, 3c, t+1]
        q, k, v = (part.reshape(n * heads, ch, -1) for part in qkv.chunk(3, dim=1))
        scale = 1.0 / math.sqrt(math.sqrt(ch))                               # applied to q and k (classifier.py:181-186)
        logits = torch.einsum("bc,bcs->bs", q[:, :, 0] * scale, k * scale)   # query = position 0 only
        weight = torch.softmax(logits, dim=-1)
        pooled = torch.einsum("bs,bcs->bc", weight, v).reshape(n, c)
        return F.linear(pooled, self.c_proj.weight[:, :, 0], self.c_proj.bias)


class ClassifierStem(nn.Module):
    """reference models/classifier.py:48-121: [N x 1 x T] -> [N x base_channels*output_mult]."""

    def __init__(
        self,
        base_channels: int = 32,
        channel_mult: Sequence[int] = DEFAULT_MULT,
        output_mult: int = 16,
        depth_mult: int = 2,
    ):
        super().__init__()
        self.base_channels = base_channels
        self.channel_mult = channel_mult
        self.output_mult = output_mult
        self.depth_mult = depth_mult
        self.out_channels = base_channels * output_mult
        self.embed_dim = emb = 4 * base_channels

        self.time_embed = TimeEmbedding(emb)
        self.time_embed_extra = nn.Sequential(_Slot("gelu"), nn.Linear(emb, emb))
        self.in_conv = nn.Conv1d(1, base_channels, 3, padding=1)
        blocks = []
        cur = base_channels
        for m in channel_mult:  # unlike the UNet, EVERY level ends with a downsampling block
            for _ in range(depth_mult):
                blocks.append(ResBlock(cur, emb, m * base_channels))
                cur = m * base_channels
            blocks.append(ResBlock(cur, emb, cur, scale_factor=0.5))
        self.blocks = nn.ModuleList(blocks)
        self.out = nn.Sequential(
            nn.Sequential(nn.GroupNorm(group_count(cur), cur), _Slot("gelu")),
            AttentionPool1d(cur, head_channels=min(cur, 64), out_channels=self.out_channels),
        )

    def conditional_embedding(self, ts: torch.Tensor, **kwargs) -> torch.Tensor:
        return time_embedding_autograd(self.time_embed, self.time_embed_extra, ts)

    def forward(self, x: torch.Tensor, ts: torch.Tensor, use_checkpoint: bool = False, **kwargs) -> torch.Tensor:
        """The stem on its own is only reached through `forward_aten` (diagnostic); `Classifier.forward` runs stem and
        head as one native program."""
        if not _use_aten() and not getattr(self, "_aten_ok", False):
            raise RuntimeError("ClassifierStem is evaluated inside Classifier.forward's libvqvs program; calling the stem "
                               "alone is only available as the ATen diagnostic (classifier.forward_aten / VQVS_GUIDANCE=aten)")
        emb = self.conditional_embedding(ts, **kwargs)
        h = _conv(x, self.in_conv)
        for blk in self.blocks:
            if use_checkpoint and torch.is_grad_enabled():
                from torch.utils.checkpoint import checkpoint

                h = checkpoint(resblock_autograd, blk, h, emb, use_reentrant=False)
            else:
                h = resblock_autograd(blk, h, emb)
        h = F.gelu(_group_norm(h, self.out[0][0]))
        return self.out[1](h)

    def load_from_predictor(self, pred) -> int:
        """reference models/classifier.py:123-130: copy the UNet's down path into the stem."""
        pairs = zip(
            [self.in_conv, self.time_embed, self.time_embed_extra, *self.blocks],
            [pred.in_conv, pred.time_embed, pred.time_embed_extra, *pred.down_blocks],
        )
        total = 0
        for dst, src in pairs:
            state = src.state_dict()
            dst.load_state_dict(state)
            total += sum(int(v.numel()) for v in state.values())
        return total


class Classifier(Savable):
    """reference models/classifier.py:18-45: stem + GELU + (zero-initialised) Linear head."""

    def __init__(self, num_labels: int, **kwargs):
        super().__init__()
        self.num_labels = num_labels
        self.stem = ClassifierStem(**kwargs)
        self.out = nn.Sequential(_Slot("gelu"), _scaled_(nn.Linear(self.stem.out_channels, num_labels), 0.0))
        self._plans = engine.PlanCache()

    def forward(self, x: torch.Tensor, ts: torch.Tensor, use_checkpoint: bool = False, **kwargs) -> torch.Tensor:
        """[N x 1 x T], [N] -> logits [N x num_labels]; differentiable w.r.t. x (native dgrad program).
        use_checkpoint trades memory for recompute in the reference; the native program keeps what it needs."""
        if _use_aten():
            return forward_aten(self, x, ts, use_checkpoint=use_checkpoint)
        from .guidance import ClassifierFunction

        engine._require_cuda(x, ts)
        return ClassifierFunction.apply(x, ts, self)

    def save_kwargs(self) -> Dict[str, Any]:
        return dict(
            num_labels=self.num_labels,
            base_channels=self.stem.base_channels,
            channel_mult=self.stem.channel_mult,
            output_mult=self.stem.output_mult,
            depth_mult=self.stem.depth_mult,
        )


def forward_aten(clf: "Classifier", x: torch.Tensor, ts: torch.Tensor, use_checkpoint: bool = False) -> torch.Tensor:
    """Diagnostic: the same parameters evaluated with torch ops under autograd (any device).  Not a fallback -- nothing in
    the package calls it unless VQVS_GUIDANCE=aten is set for an A/B timing."""
    clf.stem._aten_ok = True
    try:
        h = clf.stem(x, ts, use_checkpoint=use_checkpoint)
    finally:
        clf.stem._aten_ok = False
    head = clf.out[1]
    return F.linear(F.gelu(h), head.weight, head.bias)


class EncoderPredictor(Savable):
    """VQ-code predictor for decode-time guidance (reference models/encoder_predictor.py:15-75): a UNetPredictor with
    `bottleneck_dim` output maps, nearest sampling down to the code rate, and a 1x1 conv to `num_latents` logits.  Forward
    and the input gradient that VQVAE.decode's `enc_pred` guidance asks for (vq_vae.py:125-130) are libvqvs programs
    (guidance.PredictorFunction); the cross-entropy on the [N x D x T1] logits is left to torch (a few KB)."""

    def __init__(self, base_channels: int, downsample_rate: int, num_latents: int, bottleneck_dim: int = 64):
        super().__init__()
        self.base_channels = base_channels
        self.downsample_rate = downsample_rate
        self.num_latents = num_latents
        self.bottleneck_dim = bottleneck_dim
        self.unet = UNetPredictor(base_channels, out_channels=bottleneck_dim)
        self.out = nn.Conv1d(bottleneck_dim, num_latents, 1)
        self._plans = engine.PlanCache()

    def forward(self, x: torch.Tensor, ts: torch.Tensor, use_checkpoint: bool = False) -> torch.Tensor:
        """[N x 1 x T], [N] -> logits [N x num_latents x T // downsample_rate]."""
        from .guidance import PredictorFunction

        if x.shape[-1] % self.downsample_rate:
            raise ValueError(f"sequence length {x.shape[-1]} must be divisible by the downsample rate {self.downsample_rate}")
        return PredictorFunction.apply(x, ts, self)

    def losses(self, x: torch.Tensor, ts: torch.Tensor, targets: torch.Tensor, **kwargs) -> torch.Tensor:
        losses = F.cross_entropy(self(x, ts, **kwargs), targets, reduction="none")
        return losses.mean(-1)

    def save_kwargs(self) -> Dict[str, Any]:
        return dict(base_channels=self.base_channels, downsample_rate=self.downsample_rate, num_latents=self.num_latents,
                    bottleneck_dim=self.bottleneck_dim)
