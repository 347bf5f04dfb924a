"""Waveform VQ-VAE with a diffusion decoder: encode / decode surface of reference vq_vae.py."""

from typing import Any, Dict, Optional

import torch

from . import engine
from .diffusion_model import DiffusionModel
from .make import make_encoder
from .vq import VQ


class VQVAE(DiffusionModel):
    def __init__(self, base_channels: int, enc_name: str = "unet", cond_mult: int = 16, dictionary_size: int = 512,
                 **kwargs):
        encoder = make_encoder(enc_name=enc_name, base_channels=base_channels, cond_mult=cond_mult)
        kwargs["cond_channels"] = base_channels * cond_mult
        super().__init__(base_channels=base_channels, **kwargs)
        self.enc_name = enc_name
        self.cond_mult = cond_mult
        self.dictionary_size = dictionary_size
        self.encoder = encoder
        self.vq = VQ(self.cond_channels, dictionary_size)

    def encode(self, inputs: torch.Tensor) -> torch.Tensor:
        """[N x 1 x T] waveform -> int64 [N x T1] codes (reference vq_vae.py:82-90)."""
        with torch.no_grad():
            return self.vq(self.encoder(inputs))["idxs"]

    def decode(self, codes: torch.Tensor, labels: Optional[torch.Tensor] = None, steps: int = 100,
               progress: bool = False, constrain: bool = False, enc_pred=None, enc_pred_scale: float = 1.0,
               **kwargs) -> torch.Tensor:
        """Sample audio from codes [N x T1] (or code embeddings [N x C x T1]) and speaker labels
        (reference vq_vae.py:92-145)."""
        if len(codes.shape) == 2:
            cond_seq = self.vq.embed(codes)
        elif len(codes.shape) == 3:
            cond_seq = codes
        else:
            raise ValueError(f"unsupported codes shape: {codes.shape}")

        cond_fn = None
        if enc_pred is not None:
            targets = self.vq(cond_seq)["idxs"]

            def cond_fn(x, ts):
                with torch.enable_grad():
                    x_grad = x.detach().clone().requires_grad_(True)
                    losses = enc_pred.losses(x_grad, ts, targets) * targets.shape[-1]
                    grads = torch.autograd.grad(losses.sum(), x_grad)[0]
                return grads * enc_pred_scale * -1

        # x_T comes from the CPU generator, then moves -- exactly like the reference (:132-134)
        x_T = torch.randn(codes.shape[0], 1, codes.shape[-1] * self.encoder.downsample_rate).to(codes.device)
        bound = engine.BoundPredictor(self.predictor, cond=cond_seq, labels=labels)
        return self.diffusion.ddpm_sample(x_T, bound, steps=steps, progress=progress, constrain=constrain,
                                          cond_fn=cond_fn, **kwargs)

    def decode_uncond_guidance(self, codes, labels=None, steps: int = 100, progress: bool = False,
                               constrain: bool = False, label_scale: float = 0.0, vq_scale: float = 0.0, **kwargs):
        """Classifier-free guidance decode (reference vq_vae.py:147-220): up to three predictor
        evaluations per step at 3x batch, mixed linearly."""
        if len(codes.shape) == 2:
            cond_seq = self.vq.embed(codes)
        elif len(codes.shape) == 3:
            cond_seq = codes
        else:
            raise ValueError(f"unsupported codes shape: {codes.shape}")
        n = len(cond_seq)
        x_T = torch.randn(codes.shape[0], 1, codes.shape[-1] * self.encoder.downsample_rate).to(codes.device)
        use_vq = bool(vq_scale)
        use_label = labels is not None and bool(label_scale)

        def pred_fn(xs, ts):
            conds, labs = [cond_seq], [labels + 1]
            if use_vq:
                conds.append(torch.zeros_like(cond_seq))
                labs.append(labels + 1)
            if use_label:
                conds.append(cond_seq)
                labs.append(torch.zeros_like(labels))
            reps = len(conds)
            outs = self.predictor(torch.cat([xs] * reps), torch.cat([ts] * reps), cond=torch.cat(conds),
                                  labels=torch.cat(labs))
            base, pred, k = outs[:n], outs[:n], 1
            for on, scale in ((use_vq, vq_scale), (use_label, label_scale)):
                if on:
                    pred = pred + scale * (base - outs[k * n:(k + 1) * n])
                    k += 1
            return pred

        return self.diffusion.ddpm_sample(x_T, pred_fn, steps=steps, progress=progress, constrain=constrain, **kwargs)

    def save_kwargs(self) -> Dict[str, Any]:
        res = super().save_kwargs()
        res.update(dict(enc_name=self.enc_name, cond_mult=self.cond_mult, dictionary_size=self.dictionary_size))
        return res
