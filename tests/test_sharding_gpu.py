"""Device-side keyed noise against its numpy restatement, and N-GPU == 1-GPU sampling on hardware.

The 2-GPU case needs two devices (`gpurun --gpus 2`); it is skipped on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import rel_l2
from oracle import keyed_noise as KN
from vq_voice_swap_b200 import sharding

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


@pytest.mark.parametrize("length", [1, 6, 64, 4001])
def test_keyed_noise_matches_numpy_restatement(length):
    got = sharding.keyed_noise(99, 5, 3, 7, length, DEV).cpu().numpy().reshape(3, length)
    ref = KN.keyed_normal(99, range(5, 8), 7, length)
    # identical Philox words; float32 log / sincos differ by a few ulp between libm and the device
    assert np.abs(got - ref).max() <= 2e-5


def test_keyed_noise_sharding_invariance_and_moments():
    full = sharding.keyed_noise(3, 0, 8, 2, 64000, DEV)
    parts = torch.cat([sharding.keyed_noise(3, 0, 3, 2, 64000, DEV), sharding.keyed_noise(3, 3, 5, 2, 64000, DEV)])
    assert torch.equal(full, parts)
    assert not torch.equal(full, sharding.keyed_noise(3, 0, 8, 3, 64000, DEV))
    assert abs(float(full.mean())) < 5e-3 and abs(float(full.std()) - 1.0) < 5e-3
    assert abs(float((full[0] * full[1]).mean())) < 2e-2  # rows are independent streams
    assert sharding.keyed_noise(3, 0, 0, 2, 64, DEV).shape == (0, 1, 64)


def test_noise_fn_hook_is_used_by_the_fused_sampler():
    """ddpm_sample(noise_fn=...) with keyed noise: a batch equals its two halves sampled separately (one process)."""
    from sharded_worker import build

    model = build(DEV, num_labels=3)
    labels = (torch.arange(6) % 3).to(DEV)
    full = sharding.sample_sharded(model, 6, 3, seed=1234, length=2048, device=DEV, labels=labels)
    from vq_voice_swap_b200.engine import BoundPredictor

    halves = []
    for lo, hi in ((0, 2), (2, 6)):
        x_T = sharding.keyed_noise(1234, lo, hi - lo, -1, 2048, DEV)
        halves.append(model.diffusion.ddpm_sample(
            x_T, BoundPredictor(model.predictor, labels=labels[lo:hi]), 3,
            noise_fn=lambda i, like, lo=lo, hi=hi: sharding.keyed_noise(1234, lo, hi - lo, i, 2048, like.device)))
    assert rel_l2(torch.cat(halves).cpu(), full.cpu()) <= 1e-5  # only the atomics' summation order differs


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("total", [6, 1])  # total = 1: rank 1's shard is empty and must still join the all_gather
def test_two_gpu_sampling_equals_one_gpu(tmp_path, total):
    from sharded_worker import build

    out = str(tmp_path / "full.pt")
    steps, length = 3, 4096
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "sharded_worker.py"), out, str(total), str(steps), str(length)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    two = torch.load(out)
    model = build(DEV, num_labels=3)
    one = sharding.sample_sharded(model, total, steps, seed=1234, length=length, device=DEV, labels=torch.arange(total) % 3)
    assert two.shape == (total, 1, length)
    assert rel_l2(two, one.cpu()) <= 1e-5
