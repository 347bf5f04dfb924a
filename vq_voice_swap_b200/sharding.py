"""Batch sharding for multi-GPU sampling.

Every op on the sampling path is per-sample (GroupNorm normalises inside a sample; `constrain`
averages over time inside a sample), so a batch splits across GPUs with NO data-path collective:
rank r owns samples [lo, hi), replicas hold the same weights, and one all_gather re-assembles the
finished waveforms.  Noise is keyed by (seed, GLOBAL sample index, step) and drawn on the device
(vqvs_keyed_normal: Philox4x32-10 + Box-Muller), so the result does not depend on how many GPUs
shared the work and costs no host RNG or H2D copy.
"""

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import lib as L


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range of sample indices for `rank` (first total % world ranks get one extra)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def keyed_noise(seed: int, first_index: int, count: int, step: int, length: int, device) -> torch.Tensor:
    """[count, 1, length] standard-normal noise on `device`; row i depends only on (seed, first_index + i, step)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("keyed_noise draws on the GPU (vqvs_keyed_normal); there is no CPU fallback")
    out = torch.empty(count, 1, length, device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        L.check(L.load().vqvs_keyed_normal(out.data_ptr(), count, length, seed & 0xFFFFFFFFFFFFFFFF, first_index, step,
                                           L.stream_ptr(device)), "vqvs_keyed_normal")
    return out


def gather_samples(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """all_gather the per-rank shards (possibly of unequal or zero length) back into the [total, ...] batch."""
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


def sample_sharded(model, total: int, steps: int, seed: int, length: int = 64000, device=None,
                   labels: Optional[torch.Tensor] = None, cond: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
    """Draw `total` samples with the batch split over the ranks of the default process group.

    x_T and every step's noise are keyed by the global sample index; `labels` / `cond` (for conditional models) are given
    for the WHOLE batch and sliced per rank.  A rank whose shard is empty (total < world size) skips sampling but still
    joins the all_gather.  Returns the full batch on every rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(total, rank, world)
    device = torch.device(device or next(model.parameters()).device)
    if hi > lo:
        from .engine import BoundPredictor

        x_T = keyed_noise(seed, lo, hi - lo, -1, length, device)
        bound = {}
        if labels is not None:
            bound["labels"] = labels[lo:hi].to(device)
        if cond is not None:
            bound["cond"] = cond[lo:hi].to(device)
        predictor = BoundPredictor(model.predictor, **bound) if bound else model.predictor
        local = model.diffusion.ddpm_sample(
            x_T, predictor, steps, noise_fn=lambda i, like: keyed_noise(seed, lo, hi - lo, i, length, like.device), **kwargs)
    else:
        local = torch.empty(0, 1, length, device=device)
    return gather_samples(local, total)
