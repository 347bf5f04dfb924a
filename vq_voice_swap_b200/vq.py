"""Vector-quantisation layer (reference vq.py:73-143, 199-243) on libvqvs kernels.

Inference surface only: forward / embed / dictionary / usage_count.  The nearest-code search is
vqvs_vq_argmin (fp64-accumulated dots, the reference's fp32 expression order, first-minimum
tie-break); the training-time losses and dead-code revival are outside the sampling path.
"""

from typing import Dict

import torch
import torch.nn as nn

from . import engine
from . import lib as L


class VQ(nn.Module):
    def __init__(self, num_channels: int, num_codes: int, dead_rate: int = 100):
        super().__init__()
        self.num_channels = num_channels
        self.num_codes = num_codes
        self.dead_rate = dead_rate
        self.dictionary = nn.Parameter(torch.randn(num_codes, num_channels))
        self.register_buffer("usage_count", dead_rate * torch.ones(num_codes).long())
        self._last_batch = None

    def embed(self, idxs: torch.Tensor) -> torch.Tensor:
        """[N x ...] code indices -> [N x C x ...] embeddings (reference vq.py:98-110)."""
        engine._require_cuda(idxs, self.dictionary)
        n = idxs.shape[0]
        flat = idxs.reshape(n, -1).to(torch.int64).contiguous()
        if flat.numel():
            lo, hi = int(flat.min()), int(flat.max())
            if lo < 0 or hi >= self.num_codes:  # F.embedding (reference vq.py:108) raises the same way
                raise IndexError(f"code index out of range: [{lo}, {hi}] with a dictionary of {self.num_codes}")
        out = torch.empty(n, self.num_channels, flat.shape[1], device=idxs.device, dtype=torch.float32)
        d = engine._f32(self.dictionary)
        with torch.cuda.device(idxs.device):
            L.check(L.load().vqvs_vq_embed(flat.data_ptr(), d.data_ptr(), n, self.num_channels, flat.shape[1],
                                           self.num_codes, out.data_ptr(), L.stream_ptr(idxs.device)), "vqvs_vq_embed")
        return out.reshape(n, self.num_channels, *idxs.shape[1:])

    def forward(self, inputs: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Quantise an [N x C x ...] tensor: {"embedded", "passthrough", "idxs"} (reference vq.py:112-143)."""
        engine._require_cuda(inputs, self.dictionary)
        if inputs.shape[1] != self.num_channels:
            raise ValueError(f"expected {self.num_channels} channels, got {inputs.shape[1]}")
        n = inputs.shape[0]
        with torch.no_grad():
            x = engine._f32(inputs).reshape(n, self.num_channels, -1)
            t1 = x.shape[2]
            idxs = torch.empty(n, t1, device=x.device, dtype=torch.int64)
            d = engine._f32(self.dictionary)
            with torch.cuda.device(x.device):
                L.check(L.load().vqvs_vq_argmin(x.data_ptr(), d.data_ptr(), n, self.num_channels, t1, self.num_codes,
                                                idxs.data_ptr(), L.stream_ptr(x.device)), "vqvs_vq_argmin")
            idxs = idxs.reshape(n, *inputs.shape[2:])
            embedded = self.embed(idxs)
            if self.training:
                self._update_tracker(idxs)
                self._last_batch = x.permute(0, 2, 1).reshape(-1, self.num_channels)
        # straight-through value: embedded + (x - x.detach()) equals embedded numerically
        return {"embedded": embedded, "passthrough": embedded, "idxs": idxs}

    def _update_tracker(self, idxs: torch.Tensor):
        """usage_count semantics of reference vq.py:190-196, on the device and with a valid dtype
        (the reference's np.int crashes on numpy >= 1.24 -- SURVEY.md D6)."""
        update = torch.full_like(self.usage_count, -1)
        update[idxs.reshape(-1).unique()] = self.dead_rate
        self.usage_count.add_(update).clamp_(0, self.dead_rate)
