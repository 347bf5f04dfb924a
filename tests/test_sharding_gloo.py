"""N>1 host logic on CPU with the gloo backend (world_size 2 and 3): shard ranges, sharding-invariant
noise, and the final all_gather.  The CUDA predictor is replaced by a closed-form stand-in so the
test exercises exactly the multi-rank plumbing bench.py / sample_sharded use."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vq_voice_swap_b200 import sharding


def test_shard_ranges_cover_and_balance():
    for total in (1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_keyed_noise_is_sharding_invariant():
    full = sharding.keyed_noise(7, range(0, 6), 3, 50)
    parts = torch.cat([sharding.keyed_noise(7, range(0, 2), 3, 50), sharding.keyed_noise(7, range(2, 6), 3, 50)])
    assert torch.equal(full, parts)
    assert not torch.equal(full, sharding.keyed_noise(7, range(0, 6), 4, 50))
    assert abs(float(full.std()) - 1.0) < 0.2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy_sample(total, lo, hi, steps, seed, length):
    """A deterministic 'sampler' built from the same keyed-noise calls sample_sharded makes."""
    x = sharding.keyed_noise(seed, range(lo, hi), -1, length)
    for s in range(steps):
        x = 0.9 * x + 0.1 * sharding.keyed_noise(seed, range(lo, hi), s, length)
    return x


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(total, rank, world)
        local = _toy_sample(total, lo, hi, 3, 11, 40)
        full = sharding.gather_samples(local, total)
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 6), (3, 7)])
def test_gather_equals_single_process(tmp_path, world, total):
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ref = _toy_sample(total, 0, total, 3, 11, 40)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert got.shape == ref.shape and torch.equal(got, ref)
