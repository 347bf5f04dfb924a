"""N>1 host logic on CPU with the gloo backend (world_size 2 and 3): shard ranges, sharding-invariant
noise, and the final all_gather (including an EMPTY shard).  The CUDA predictor is replaced by a
closed-form stand-in and the device noise kernel by its numpy restatement (oracle/keyed_noise.py), so
the test exercises exactly the multi-rank plumbing bench.py / sample_sharded use."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import keyed_noise as KN
from vq_voice_swap_b200 import sharding


def _noise(seed, lo, hi, step, length):
    return torch.from_numpy(KN.keyed_normal(seed, range(lo, hi), step, length)).reshape(hi - lo, 1, length)


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
    full = _noise(7, 0, 6, 3, 50)
    parts = torch.cat([_noise(7, 0, 2, 3, 50), _noise(7, 2, 6, 3, 50)])
    assert torch.equal(full, parts)
    assert not torch.equal(full, _noise(7, 0, 6, 4, 50))
    assert not torch.equal(full, _noise(8, 0, 6, 3, 50))
    big = _noise(1, 0, 4, -1, 20000)
    assert abs(float(big.mean())) < 0.02 and abs(float(big.std()) - 1.0) < 0.02


def test_keyed_noise_refuses_cpu():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sharding.keyed_noise(1, 0, 2, 0, 16, "cpu")


def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10."""
    for ctr, key, want in [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]:
        got = KN.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy_sample(total, lo, hi, steps, seed, length):
    """A deterministic 'sampler' built from the same keyed-noise calls sample_sharded makes."""
    x = _noise(seed, lo, hi, -1, length)
    for s in range(steps):
        x = 0.9 * x + 0.1 * _noise(seed, lo, hi, s, length)
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


@pytest.mark.parametrize("world,total", [(2, 6), (3, 7), (3, 2)])  # (3, 2): the last rank's shard is empty
def test_gather_equals_single_process(tmp_path, world, total):
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ref = _toy_sample(total, 0, total, 3, 11, 40)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert got.shape == ref.shape and torch.equal(got, ref)
