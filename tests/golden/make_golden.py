"""
Generate the golden vectors in this directory from the LIVE, UNMODIFIED reference.

Run in the build container only (the reference is mounted at /root/reference
there and exists nowhere else):

    python tests/golden/make_golden.py            (everything)
    python tests/golden/make_golden.py classifier (classifier vectors + key table only)

The script puts /root/reference FIRST on sys.path so that `vq_voice_swap` resolves to the reference and not to
this repository's drop-in namespace of the same name.

Inputs and weights are pure functions of their names
(vq_voice_swap_b200.synth), so only the reference's OUTPUTS are stored.  The
tests replay these files against oracle/ (CPU) and against the CUDA path (GPU).
"""

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(os.path.dirname(HERE))]
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
sys.path.append(os.path.dirname(os.path.dirname(HERE)))  # for vq_voice_swap_b200.synth only

import vq_voice_swap  # noqa: E402  (must resolve to the reference)

assert vq_voice_swap.__file__.startswith("/root/reference"), vq_voice_swap.__file__

from vq_voice_swap.diffusion import Diffusion, make_schedule  # noqa: E402
from vq_voice_swap.diffusion_model import DiffusionModel  # noqa: E402
from vq_voice_swap.models.classifier import Classifier  # noqa: E402
from vq_voice_swap.models.unet import ResBlock, UNetEncoder  # noqa: E402
from vq_voice_swap.vq import VQ  # noqa: E402
from vq_voice_swap.vq_vae import VQVAE  # noqa: E402

from vq_voice_swap_b200 import synth  # noqa: E402

torch.set_grad_enabled(False)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name), **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrays.items()})


# Shared with the tests: the case tables live in cases.py so both sides agree.
from cases import DDPM_CASES, RESBLOCK_CASES  # noqa: E402


def resblocks():
    out = {}
    for name, kw in RESBLOCK_CASES.items():
        blk = ResBlock(**kw["ctor"]).eval()
        synth.load_synth(blk, tag=f"rb/{name}")
        x = synth.normal(f"rb/{name}/x", (kw["batch"], kw["ctor"]["channels"], kw["t"]))
        emb = None
        if kw["ctor"].get("emb_channels"):
            emb = synth.normal(f"rb/{name}/emb", (kw["batch"], kw["ctor"]["emb_channels"]))
        out[name] = blk(x, emb).numpy()
    save("resblocks.npz", **out)


def unet_small():
    out = {}
    m = DiffusionModel("unet", 16).eval()
    synth.load_synth(m, tag="unet16")
    x = synth.normal("unet16/x", (2, 1, 512))
    ts = torch.tensor([0.9, 0.3])
    out["uncond"] = m.predictor(x, ts).numpy()

    m = DiffusionModel("unet", 16, num_labels=5, cond_channels=48).eval()
    synth.load_synth(m, tag="unet16c")
    cond = synth.normal("unet16c/cond", (2, 48, 2))
    labels = torch.tensor([4, 1])
    out["cond"] = m.predictor(x, ts, cond=cond, labels=labels).numpy()

    enc = UNetEncoder(16, out_channels=48).eval()
    synth.load_synth(enc, tag="enc16")
    out["encoder"] = enc(x).numpy()
    save("unet_bc16.npz", **out)


def vq_cases():
    vq = VQ(48, 96).eval()
    synth.load_synth(vq, tag="vq")
    x = synth.normal("vq/x", (3, 48, 40))
    res = vq(x)
    # exact duplicate dictionary rows -> first-index tie-break must be visible
    vq2 = VQ(16, 32).eval()
    sd = synth.synth_state_dict(synth.shapes_of(vq2), tag="vq2")
    sd["dictionary"][20] = sd["dictionary"][7]
    sd["dictionary"][31] = sd["dictionary"][7]
    vq2.load_state_dict(sd)
    x2 = sd["dictionary"][synth.integers("vq2/pick", (2, 50), 32)].permute(0, 2, 1).contiguous()
    x2 = x2 + 0.01 * synth.normal("vq2/jitter", x2.shape)
    save(
        "vq.npz",
        idxs=res["idxs"].numpy(),
        embedded=res["embedded"].numpy(),
        idxs_ties=vq2(x2)["idxs"].numpy(),
        embed_from_idx=vq.embed(synth.integers("vq/codes", (2, 7), 96)).numpy(),
    )


class _Inject:
    """Replace torch.randn / randn_like by deterministic synthetic draws."""

    def __init__(self, tag):
        self.tag, self.n = tag, 0

    def __enter__(self):
        self._randn, self._like = torch.randn, torch.randn_like
        torch.randn = lambda *shape, **kw: synth.normal(f"{self.tag}/x_T", shape if not isinstance(shape[0], (tuple, list)) else shape[0])

        def like(x, **kw):
            t = synth.normal(f"{self.tag}/noise{self.n}", x.shape)
            self.n += 1
            return t.to(x)

        torch.randn_like = like
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._randn, self._like


def ddpm():
    out = {}
    for name, kw in DDPM_CASES.items():
        diff = Diffusion(make_schedule(kw["schedule"]))
        x_t = synth.normal(f"ddpm/{name}/x", (3, 1, 96), std=kw.get("x_std", 1.0))
        eps = synth.normal(f"ddpm/{name}/eps", (3, 1, 96))
        noise = synth.normal(f"ddpm/{name}/noise", (3, 1, 96))
        ts = torch.tensor(kw["ts"], dtype=torch.float32)
        cond_fn = (lambda x, t: torch.sin(x) * t[:, None, None]) if kw.get("cond_fn") else None
        out[name] = diff.ddpm_previous(
            x_t, ts, kw["step"], eps, noise=noise, sigma_large=kw.get("sigma_large", False),
            constrain=kw.get("constrain", False), cond_fn=cond_fn,
        ).numpy()

    # whole sampler loop on a closed-form predictor; with and without sample-time schedule
    toy = lambda x, ts: 0.7 * x * ts[:, None, None] + 0.1
    for name, sched, constrain in [("loop_plain", None, False), ("loop_sq", lambda t: t ** 2, True)]:
        diff = Diffusion(make_schedule("exp"))
        x_T = synth.normal(f"ddpm/{name}/x_T", (2, 1, 64))
        with _Inject(f"ddpm/{name}"):
            out[name] = diff.ddpm_sample(x_T, toy, 6, constrain=constrain, schedule=sched).numpy()
    save("ddpm.npz", **out)


def vqvae_small():
    m = VQVAE(base_channels=16, num_labels=3, cond_mult=3, dictionary_size=64, pred_name="unet").eval()
    synth.load_synth(m, tag="vqvae16")
    w = synth.normal("vqvae16/wave", (2, 1, 512)).clamp(-1, 1)
    codes = m.encode(w)
    enc_out = m.encoder(w)
    labels = torch.tensor([2, 0])
    with _Inject("vqvae16/decode"):
        audio = m.decode(codes, labels, steps=3, constrain=True)
    save("vqvae_bc16.npz", codes=codes.numpy(), encoder_out=enc_out.numpy(), audio=audio.numpy())


def vqvae_uncond():
    """decode_uncond_guidance (vq_vae.py:147-220).  The reference always repeats x three times (:191-193), so it only works
    with BOTH guidance scales non-zero; label 0 is the unconditional label, real labels are offset by one."""
    m = VQVAE(base_channels=16, num_labels=4, cond_mult=3, dictionary_size=64, pred_name="unet").eval()
    synth.load_synth(m, tag="vqvae16u")
    codes = synth.integers("vqvae16u/codes", (2, 2), 64)
    labels = torch.tensor([2, 0])
    with _Inject("vqvae16u/decode"):
        audio = m.decode_uncond_guidance(codes, labels, steps=3, constrain=True, label_scale=1.3, vq_scale=0.7)
    save("vqvae_uncond_bc16.npz", audio=audio.numpy())


def conv_mfcc():
    """ConvMFCCEncoder (models/conv_encoder.py) version 1, mu-law and linear input; only the conv stack's parameters are
    re-randomised (the transform's window / filter bank / DCT buffers keep torchaudio's values)."""
    from vq_voice_swap.models.conv_encoder import ConvMFCCEncoder

    out = {}
    for name, ulaw in (("ulaw", True), ("linear", False)):
        m = ConvMFCCEncoder(base_channels=8, out_channels=32, input_ulaw=ulaw).eval()
        sd = m.state_dict()
        shapes = {k: (tuple(v.shape), v.dtype) for k, v in sd.items() if k.startswith("blocks.")}
        sd.update(synth.synth_state_dict(shapes, tag="convmfcc"))
        m.load_state_dict(sd)
        x = (0.4 * synth.normal("convmfcc/x", (2, 1, 6400))).clamp(-1, 1)
        out[name] = m(x).numpy()
        out[name + "_mfcc"] = m.mfcc(x[:, 0] if not ulaw else x[:, 0].sign() * (1 / 255.0) * (256.0 ** x[:, 0].abs() - 1)).numpy()
    out["keys"] = np.array([f"{k}|{tuple(v.shape)}" for k, v in m.state_dict().items()])
    save("conv_mfcc.npz", **out)


def encoder_predictor():
    """EncoderPredictor (models/encoder_predictor.py): logits and the guidance gradient VQVAE.decode asks for
    (vq_vae.py:125-130: d/dx sum(losses * T1))."""
    from vq_voice_swap.models import EncoderPredictor

    torch.set_grad_enabled(True)
    m = EncoderPredictor(base_channels=16, downsample_rate=256, num_latents=32, bottleneck_dim=16).eval()
    synth.load_synth(m, tag="encpred16")
    x = synth.normal("encpred16/x", (2, 1, 1024)).requires_grad_()
    ts = torch.tensor([0.7, 0.3])
    targets = synth.integers("encpred16/targets", (2, 4), 32)
    logits = m(x, ts)
    losses = m.losses(x, ts, targets) * targets.shape[-1]
    grad = torch.autograd.grad(losses.sum(), x)[0]
    torch.set_grad_enabled(False)
    save("encoder_predictor16.npz", logits=logits.detach().numpy(), grad=grad.numpy(), losses=losses.detach().numpy(),
         keys=np.array([f"{k}|{tuple(v.shape)}" for k, v in m.state_dict().items()]))


def classifier_small():
    """Classifier logits, the guidance gradient of sample_diffusion.py:34-42, and one guided ddpm_previous."""
    import torch.nn.functional as F

    torch.set_grad_enabled(True)
    clf = Classifier(num_labels=7, base_channels=16).eval()
    synth.load_synth(clf, tag="clf16")
    x = synth.normal("clf16/x", (2, 1, 1024))
    ts = torch.tensor([0.8, 0.25])
    labels = torch.tensor([3, 6])
    logits = clf(x, ts)

    def cond_fn(xx, tt):
        with torch.enable_grad():
            xx = xx.detach().clone().requires_grad_()
            logp = F.log_softmax(clf(xx, tt), dim=-1)
            return torch.autograd.grad(logp[range(len(xx)), labels].sum(), xx)[0].detach() * 2.5

    grad = cond_fn(x, ts)
    diff = Diffusion(make_schedule("exp"))
    eps = synth.normal("clf16/eps", (2, 1, 1024))
    noise = synth.normal("clf16/noise", (2, 1, 1024))
    prev = diff.ddpm_previous(x, ts, 0.02, eps, noise=noise, cond_fn=cond_fn)
    torch.set_grad_enabled(False)
    save("classifier_bc16.npz", logits=logits.detach().numpy(), grad=grad.numpy(), prev=prev.numpy())


def keys():
    specs = {
        "diffusion_unet32": DiffusionModel("unet", 32),
        "diffusion_unet16_cond": DiffusionModel("unet", 16, num_labels=5, cond_channels=48),
        "diffusion_unet16_dropout": DiffusionModel("unet", 16, dropout=0.1),
        "vqvae_unet32": VQVAE(base_channels=32, pred_name="unet", num_labels=8),
        "diffusion_unet16": DiffusionModel("unet", 16),
        "vqvae16": VQVAE(base_channels=16, num_labels=3, cond_mult=3, dictionary_size=64, pred_name="unet"),
        "classifier16": Classifier(num_labels=7, base_channels=16),
        "classifier32": Classifier(num_labels=100, base_channels=32),
    }
    table = {
        n: [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()] for n, m in specs.items()
    }
    table["save_kwargs"] = {n: m.save_kwargs() for n, m in specs.items()}
    table["encoder16"] = [[k, list(v.shape), str(v.dtype)] for k, v in UNetEncoder(16, out_channels=48).state_dict().items()]
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(table, f)
    print("wrote state_dict_keys.json", {k: len(v) for k, v in table.items()})


if __name__ == "__main__":
    if sys.argv[1:] == ["keys"]:
        keys()
        sys.exit(0)
    if sys.argv[1:] == ["classifier"]:
        classifier_small()
        keys()
        sys.exit(0)
    if sys.argv[1:] == ["uncond"]:
        vqvae_uncond()
        sys.exit(0)
    if sys.argv[1:] == ["encpred"]:
        encoder_predictor()
        sys.exit(0)
    if sys.argv[1:] == ["mfcc"]:
        conv_mfcc()
        sys.exit(0)
    resblocks()
    unet_small()
    vq_cases()
    ddpm()
    vqvae_small()
    vqvae_uncond()
    conv_mfcc()
    encoder_predictor()
    classifier_small()
    keys()
