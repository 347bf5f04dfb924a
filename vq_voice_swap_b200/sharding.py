"""Batch sharding for multi-GPU sampling.

Every op on the sampling path is per-sample (GroupNorm normalises inside a sample; `constrain`
averages over time inside a sample), so a batch splits across GPUs with NO data-path collective:
rank r owns samples [lo, hi), replicas hold the same weights, and one all_gather re-assembles the
finished waveforms.  Noise is keyed by (seed, GLOBAL sample index, step), so the result does not
depend on how many GPUs shared the work.
"""

import hashlib
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range of sample indices for `rank` (first total % world ranks get one extra)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _seed_for(seed: int, index: int, step: int) -> int:
    h = hashlib.blake2b(f"{seed}/{index}/{step}".encode(), digest_size=8).digest()
    return int.from_bytes(h, "little") & 0x7FFFFFFFFFFFFFFF


def keyed_noise(seed: int, indices: range, step: int, length: int, device="cpu") -> torch.Tensor:
    """[len(indices), 1, length] standard normal noise; row i depends only on (seed, indices[i], step)."""
    rows = []
    for idx in indices:
        gen = torch.Generator(device="cpu")
        gen.manual_seed(_seed_for(seed, idx, step))
        rows.append(torch.randn(1, length, generator=gen))
    out = torch.stack(rows) if rows else torch.empty(0, 1, length)
    return out.to(device)


def gather_samples(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """all_gather the per-rank shards (possibly of unequal length) back into the [total, ...] batch."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] < width:
        padded = torch.cat([local, local.new_zeros(width - local.shape[0], *local.shape[1:])])
    parts: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])


def sample_sharded(model, total: int, steps: int, seed: int, length: int = 64000, device=None, **kwargs) -> torch.Tensor:
    """Draw `total` samples with the batch split over the ranks of the default process group.

    x_T and every step's noise are keyed by the global sample index; returns the full batch on every rank.
    """
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(total, rank, world)
    device = device or next(model.parameters()).device
    x_T = keyed_noise(seed, range(lo, hi), -1, length, device)
    step = [0]
    orig = torch.randn_like

    def keyed_like(x, **kw):  # the sampler asks for one noise tensor per step (reference diffusion.py:62-63)
        t = keyed_noise(seed, range(lo, hi), step[0], length, x.device).to(x.dtype)
        step[0] += 1
        return t

    torch.randn_like = keyed_like
    try:
        local = model.diffusion.ddpm_sample(x_T, model.predictor, steps, **kwargs)
    finally:
        torch.randn_like = orig
    return gather_samples(local, total)
