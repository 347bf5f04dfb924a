"""Pin oracle/hotpath.py to the outputs of the live reference (tests/golden)."""

import numpy as np
import pytest
import torch

from cases import DDPM_CASES, RESBLOCK_CASES
from helpers import model_sd, rel_l2, resblock_case
from oracle import hotpath as O
from vq_voice_swap_b200 import synth

TOL = 1e-6  # same ATen CPU kernels on both sides; expected bit-identical


@pytest.mark.parametrize("name", sorted(RESBLOCK_CASES))
def test_resblock(golden, name):
    kw = RESBLOCK_CASES[name]
    sd, x, emb = resblock_case(name, kw)
    ctor = kw["ctor"]
    y = O.resblock(x, emb, sd, "", scale_factor=ctor.get("scale_factor", 1.0), dilation=ctor.get("dilation", 2))
    assert rel_l2(y, golden("resblocks.npz")[name]) <= TOL


def test_unet_predictor_uncond(golden):
    sd = model_sd("diffusion_unet16", "unet16")
    x = synth.normal("unet16/x", (2, 1, 512))
    y = O.unet_predictor(sd, x, torch.tensor([0.9, 0.3]))
    assert rel_l2(y, golden("unet_bc16.npz")["uncond"]) <= TOL


def test_unet_predictor_cond(golden):
    sd = model_sd("diffusion_unet16_cond", "unet16c")
    x = synth.normal("unet16/x", (2, 1, 512))
    cond = synth.normal("unet16c/cond", (2, 48, 2))
    y = O.unet_predictor(sd, x, torch.tensor([0.9, 0.3]), cond=cond, labels=torch.tensor([4, 1]))
    assert rel_l2(y, golden("unet_bc16.npz")["cond"]) <= TOL


def test_unet_encoder(golden):
    sd = model_sd("encoder16", "enc16")
    x = synth.normal("unet16/x", (2, 1, 512))
    y = O.unet_encoder(sd, x, prefix="")
    assert rel_l2(y, golden("unet_bc16.npz")["encoder"]) <= TOL


def test_vq(golden):
    g = golden("vq.npz")
    d = synth.normal("vq/dictionary", (96, 48))
    x = synth.normal("vq/x", (3, 48, 40))
    idx = O.vq_encode(d, x)
    assert np.array_equal(idx.numpy(), g["idxs"])
    assert np.array_equal(O.vq_embed(d, idx).numpy(), g["embedded"])
    codes = synth.integers("vq/codes", (2, 7), 96)
    assert np.array_equal(O.vq_embed(d, codes).numpy(), g["embed_from_idx"])


def test_vq_ties_first_index(golden):
    d = synth.normal("vq2/dictionary", (32, 16))
    d[20] = d[7]
    d[31] = d[7]
    x = d[synth.integers("vq2/pick", (2, 50), 32)].permute(0, 2, 1).contiguous()
    x = x + 0.01 * synth.normal("vq2/jitter", x.shape)
    idx = O.vq_encode(d, x).numpy()
    assert np.array_equal(idx, golden("vq.npz")["idxs_ties"])
    assert not np.isin(idx, [20, 31]).any()  # duplicates of row 7 never win


@pytest.mark.parametrize("name", sorted(DDPM_CASES))
def test_ddpm_previous(golden, name):
    kw = DDPM_CASES[name]
    x_t = synth.normal(f"ddpm/{name}/x", (3, 1, 96), std=kw.get("x_std", 1.0))
    eps = synth.normal(f"ddpm/{name}/eps", (3, 1, 96))
    noise = synth.normal(f"ddpm/{name}/noise", (3, 1, 96))
    cond_fn = (lambda x, t: torch.sin(x) * t[:, None, None]) if kw.get("cond_fn") else None
    y = O.ddpm_previous(
        O.make_alpha_bar(kw["schedule"]), x_t, torch.tensor(kw["ts"]), kw["step"], eps, noise,
        sigma_large=kw.get("sigma_large", False), constrain=kw.get("constrain", False), cond_fn=cond_fn,
    )
    assert rel_l2(y, golden("ddpm.npz")[name]) <= TOL


@pytest.mark.parametrize("name,sched,constrain", [("loop_plain", None, False), ("loop_sq", lambda t: t ** 2, True)])
def test_ddpm_sample_loop(golden, name, sched, constrain):
    toy = lambda x, ts: 0.7 * x * ts[:, None, None] + 0.1
    x_T = synth.normal(f"ddpm/{name}/x_T", (2, 1, 64))
    noises = [synth.normal(f"ddpm/{name}/noise{i}", (2, 1, 64)) for i in range(5)]
    y = O.ddpm_sample(O.make_alpha_bar("exp"), x_T, toy, 6, noises, constrain=constrain, schedule=sched)
    assert rel_l2(y, golden("ddpm.npz")[name]) <= TOL


def test_vqvae_roundtrip(golden):
    g = golden("vqvae_bc16.npz")
    sd = model_sd("vqvae16", "vqvae16")
    w = synth.normal("vqvae16/wave", (2, 1, 512)).clamp(-1, 1)
    assert rel_l2(O.unet_encoder(sd, w), g["encoder_out"]) <= TOL
    codes = O.vqvae_encode(sd, w)
    assert np.array_equal(codes.numpy(), g["codes"])
    x_T = synth.normal("vqvae16/decode/x_T", (2, 1, 512))
    noises = [synth.normal(f"vqvae16/decode/noise{i}", (2, 1, 512)) for i in range(2)]
    audio = O.vqvae_decode(sd, "exp", codes, torch.tensor([2, 0]), 3, x_T, noises, constrain=True)
    assert rel_l2(audio, g["audio"]) <= 1e-5


def test_classifier_logits_gradient_and_guided_step(golden):
    """Classifier guidance (config 5): logits, d log p / d x and one guided ddpm_previous vs the live reference."""
    g = golden("classifier_bc16.npz")
    sd = model_sd("classifier16", "clf16")
    x = synth.normal("clf16/x", (2, 1, 1024))
    ts = torch.tensor([0.8, 0.25])
    labels = torch.tensor([3, 6])
    assert rel_l2(O.classifier_logits(sd, x, ts).detach(), g["logits"]) <= TOL
    cond_fn = O.classifier_cond_fn(sd, labels, scale=2.5)
    assert rel_l2(cond_fn(x, ts), g["grad"]) <= 1e-5
    eps = synth.normal("clf16/eps", (2, 1, 1024))
    noise = synth.normal("clf16/noise", (2, 1, 1024))
    prev = O.ddpm_previous(O.make_alpha_bar("exp"), x, ts, 0.02, eps, noise, cond_fn=cond_fn)
    assert rel_l2(prev, g["prev"]) <= 1e-5
