"""UNet predictor / encoder with the reference's constructor arguments, state-dict
layout and call signatures (reference models/unet.py), executed by libvqvs.

The nn.Module tree below exists to OWN PARAMETERS under the reference's names
(``down_blocks.3.pre_cond.2.weight`` ...), created in the reference's order so a
given torch seed initialises both identically.  No torch op runs in forward():
the call is compiled once per (batch, length) into a static launch program
(engine.py) of fused sm_100a kernels and replayed with one C call.
"""

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import engine
from .base import Encoder, Predictor

DEFAULT_MULT = (1, 1, 2, 2, 2, 4, 4, 8, 8)


def group_count(ch: int) -> int:
    """GroupNorm group count rule of reference models/unet.py:345-349."""
    g = 32
    while ch % g:
        g //= 2
    return g


class _Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices equal to the reference's
    (GELU / Resize / Dropout positions)."""

    def __init__(self, what: str = ""):
        super().__init__()
        self.what = what

    def extra_repr(self) -> str:
        return self.what

    def forward(self, *a, **k):
        raise RuntimeError("placeholder layer: the enclosing block runs as one fused CUDA program")


def _scaled_(module: nn.Module, s: float) -> nn.Module:
    with torch.no_grad():
        for p in module.parameters():
            p.mul_(s)
    return module


class TimeEmbedding(nn.Module):
    """Owns ``proj`` (reference models/wavegrad.py:352-357); evaluated inside vqvs_time_embed."""

    def __init__(self, channels: int):
        super().__init__()
        if channels % 2:
            raise AssertionError(f"channels {channels} should be divisible by two")
        self.channels = channels
        self.proj = nn.Linear(channels, channels)

    def forward(self, t):
        raise RuntimeError("TimeEmbedding is fused into the predictor program")


class ResBlock(nn.Module):
    """reference models/unet.py:248-316 -- GN/GELU/(resize)/conv3/GN/FiLM/GELU/dilated conv3 + skip."""

    def __init__(
        self,
        channels: int,
        emb_channels: Optional[int] = None,
        out_channels: Optional[int] = None,
        scale_factor: float = 1.0,
        dilation: int = 2,
        dropout: float = 0.0,
    ):
        super().__init__()
        self.channels = channels
        self.emb_channels = emb_channels
        self.out_channels = out_channels or channels
        self.scale_factor = scale_factor
        self.dilation = dilation
        self.dropout = dropout
        c_in, c_out = self.channels, self.out_channels

        projection = nn.Conv1d(c_in, c_out, 1) if c_in != c_out else _Slot("identity")
        self.skip = nn.Sequential(_Slot(f"resize x{scale_factor}"), projection)
        if emb_channels:
            self.cond_layers = nn.Sequential(_Slot("gelu"), _scaled_(nn.Linear(emb_channels, 2 * c_out), 0.1))
        self.pre_cond = nn.Sequential(
            nn.Sequential(nn.GroupNorm(group_count(c_in), c_in), _Slot("gelu")),
            _Slot(f"resize x{scale_factor}"),
            nn.Conv1d(c_in, c_out, 3, padding=1),
            nn.GroupNorm(group_count(c_out), c_out),
        )
        tail = _scaled_(nn.Conv1d(c_out, c_out, 3, padding=dilation, dilation=dilation), 0.0)
        slots = [_Slot("gelu")] + ([_Slot(f"dropout p={dropout} (inference: identity)")] if dropout else [])
        self.post_cond = nn.Sequential(*slots, tail)
        self._plans = engine.PlanCache()

    @property
    def resize_mode(self) -> int:
        return engine.resize_mode(self.scale_factor)

    def forward(self, x: torch.Tensor, cond: Optional[torch.Tensor] = None) -> torch.Tensor:
        if bool(self.emb_channels) != (cond is not None):
            raise AssertionError("must provide an embedding if and only if the block has FiLM layers")
        return engine.run_single_block(self, x, cond)


def _levels(base: int, mult: Sequence[int], depth_mult: int, emb: Optional[int], dropout: float):
    """Down path shared by predictor and encoder: yields (block, is_downsample)."""
    blocks, widths = [], [base]
    cur = base
    for depth, m in enumerate(mult):
        for _ in range(depth_mult):
            blocks.append(ResBlock(cur, emb, m * base, dropout=dropout))
            cur = m * base
            widths.append(cur)
        if depth != len(mult) - 1:
            blocks.append(ResBlock(cur, emb, scale_factor=0.5, dropout=dropout))
            widths.append(cur)
    return blocks, widths, cur


class UNetPredictor(Predictor):
    """reference models/unet.py:16-184."""

    def __init__(
        self,
        base_channels: int,
        channel_mult: Tuple[int] = DEFAULT_MULT,
        middle_dilations: Tuple[int] = (4, 8, 16, 32),
        depth_mult: int = 2,
        cond_channels: Optional[int] = None,
        num_labels: Optional[int] = None,
        in_channels: int = 1,
        out_channels: int = 1,
        dropout: float = 0.0,
    ):
        super().__init__()
        self.base_channels = base_channels
        self.channel_mult = tuple(channel_mult)
        self.middle_dilations = tuple(middle_dilations)
        self.depth_mult = depth_mult
        self.cond_channels = cond_channels
        self.num_labels = num_labels
        self.in_channels = in_channels
        self.out_channels = out_channels
        emb = 4 * base_channels

        self.time_embed = TimeEmbedding(emb)
        self.time_embed_extra = nn.Sequential(_Slot("gelu"), nn.Linear(emb, emb))
        if num_labels is not None:
            self.class_embed = nn.Embedding(num_labels, emb)
        if cond_channels is not None:
            self.cond_proj = nn.Conv1d(cond_channels, base_channels, 3, padding=1)
        self.in_conv = nn.Conv1d(in_channels, base_channels, 3, padding=1)

        down, widths, cur = _levels(base_channels, self.channel_mult, depth_mult, emb, dropout)
        self.down_blocks = nn.ModuleList(down)
        self.middle_blocks = nn.ModuleList(
            [ResBlock(cur, emb, dilation=d, dropout=dropout) for d in self.middle_dilations]
        )
        up: List[ResBlock] = []
        for depth in reversed(range(len(self.channel_mult))):
            width = self.channel_mult[depth] * base_channels
            for _ in range(depth_mult + 1):
                up.append(ResBlock(cur + widths.pop(), emb, width, dropout=dropout))
                cur = width
            if depth:
                up.append(ResBlock(cur, emb, scale_factor=2.0, dropout=dropout))
        self.up_blocks = nn.ModuleList(up)
        self.out = nn.Sequential(
            nn.Sequential(nn.GroupNorm(group_count(base_channels), base_channels), _Slot("gelu")),
            nn.Conv1d(base_channels, out_channels, 3, padding=1),
        )
        self._plans = engine.PlanCache()

    def forward(
        self,
        x: torch.Tensor,
        ts: torch.Tensor,
        cond: Optional[torch.Tensor] = None,
        labels: Optional[torch.Tensor] = None,
        use_checkpoint: bool = False,
    ) -> torch.Tensor:
        assert (labels is None) == (
            self.num_labels is None
        ), "must provide labels if and only if model is class conditional"
        assert (cond is None) == (
            self.cond_channels is None
        ), "must provide cond sequence if and only if model is conditional"
        # use_checkpoint trades memory for recompute in training; inference ignores it.
        return engine.predictor_forward(self, x, ts, cond, labels)

    def add_labels(self, n: int, end: bool = True):
        assert self.num_labels is not None
        old = self.class_embed.weight.detach()
        count = self.num_labels
        self.num_labels += n
        self.class_embed = nn.Embedding(self.num_labels, old.shape[-1]).to(old.device)
        with torch.no_grad():
            (self.class_embed.weight[:count] if end else self.class_embed.weight[n:]).copy_(old)
        self._plans.clear()

    def label_parameters(self) -> List[nn.Parameter]:
        assert self.num_labels is not None
        return list(self.class_embed.parameters())

    @property
    def downsample_rate(self) -> int:
        return 2 ** (len(self.channel_mult) - 1)


class UNetEncoder(Encoder):
    """reference models/unet.py:187-245 -- the down path without FiLM plus GN/GELU/conv3 head."""

    def __init__(
        self,
        base_channels: int,
        channel_mult: Tuple[int] = DEFAULT_MULT,
        out_dilations: Tuple[int] = (),
        depth_mult: int = 2,
        in_channels: int = 1,
        out_channels: int = 512,
    ):
        super().__init__()
        self.base_channels = base_channels
        self.channel_mult = tuple(channel_mult)
        self.out_dilations = tuple(out_dilations)
        self.depth_mult = depth_mult
        self.in_channels = in_channels
        self.out_channels = out_channels

        self.in_conv = nn.Conv1d(in_channels, base_channels, 3, padding=1)
        blocks, _, cur = _levels(base_channels, self.channel_mult, depth_mult, None, 0.0)
        blocks += [ResBlock(cur, dilation=d) for d in self.out_dilations]
        self.blocks = nn.ModuleList(blocks)
        self.out = nn.Sequential(
            nn.Sequential(nn.GroupNorm(group_count(cur), cur), _Slot("gelu")),
            nn.Conv1d(cur, out_channels, 3, padding=1),
        )
        self._plans = engine.PlanCache()

    def forward(self, x: torch.Tensor, use_checkpoint: bool = False) -> torch.Tensor:
        return engine.encoder_forward(self, x)

    @property
    def downsample_rate(self) -> int:
        return 2 ** (len(self.channel_mult) - 1)
