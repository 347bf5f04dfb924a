"""__graft_entry__.smoke(): one small hot-path invocation on cuda:0, checked against the oracle."""

import os
import sys

import torch


def run() -> None:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import hotpath as O  # checker only
    from vq_voice_swap_b200 import lib, synth
    from vq_voice_swap_b200.diffusion_model import DiffusionModel

    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device; vq_voice_swap_b200 has no CPU fallback")
    lib.load()
    dev = torch.device("cuda:0")
    model = DiffusionModel("unet", 32)
    sd = synth.synth_state_dict(synth.shapes_of(model), tag="smoke32")
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    x = synth.normal("smoke/x", (2, 1, 2048))
    ts = torch.tensor([0.7, 0.2])

    eps = model.predictor(x.to(dev), ts.to(dev))
    ref = O.unet_predictor(sd, x, ts)
    err = float((eps.cpu() - ref).norm() / ref.norm())
    plan = next(iter(model.predictor._plans.items.values()))
    kinds = [k for k, _ in plan.descs]
    print(f"smoke: unet32 forward [2,1,2048] rel_l2 vs oracle = {err:.3e}; "
          f"{plan.n_launch} kernels, {kinds.count(lib.OP_CONV_UMMA)} tcgen05 convs, {kinds.count(lib.OP_CONV_SIMT)} simt convs")
    if not err <= 1e-3:
        raise RuntimeError(f"smoke parity failed: rel_l2 {err:.3e} > 1e-3")

    # two fused DDPM steps (x_{t-1} written by the UNet's last kernel)
    torch.manual_seed(0)
    out = model.diffusion.ddpm_sample(x.to(dev), model.predictor, 2)
    if not torch.isfinite(out).all():
        raise RuntimeError("smoke: non-finite sample")
    torch.cuda.synchronize()
    print("smoke: ok")


if __name__ == "__main__":
    run()
