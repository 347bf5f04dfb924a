import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests")); sys.path.insert(0, os.path.join(os.getcwd(), "tests/golden"))
from helpers import rel_l2
from oracle import hotpath as O
from vq_voice_swap_b200 import synth
from vq_voice_swap_b200.diffusion_model import DiffusionModel
os.environ["VQVS_BACKEND"] = "umma"
m = DiffusionModel("unet", 64)
sd = synth.synth_state_dict(synth.shapes_of(m), tag="full64"); m.load_state_dict(sd); m = m.cuda().eval()
x = synth.normal("full64/x", (1, 1, 64000)); ts = torch.tensor([0.62])
ref = O.unet_predictor(sd, x, ts)
got = m.predictor(x.cuda(), ts.cuda()).cpu()
print("unet64 forward, T=64000, batch 1: rel_l2 vs CPU oracle = %.3e, max abs = %.3e (ref std %.3f)" % (rel_l2(got, ref), float((got - ref).abs().max()), float(ref.std())))
# N-step sampler with injected noise vs oracle (argv: step counts, default 4)
for steps in [int(a) for a in sys.argv[1:]] or [4]:
    x_T = synth.normal("drift/x_T", (1, 1, 64000)); noises = [synth.normal(f"drift/n{i}", x_T.shape) for i in range(steps)]
    it = iter(noises); orig = torch.randn_like
    torch.randn_like = lambda t, **k: next(it).to(t)
    y = m.diffusion.ddpm_sample(x_T.cuda(), m.predictor, steps).cpu()
    torch.randn_like = orig
    r = O.ddpm_sample(O.make_alpha_bar("exp"), x_T, lambda a, b: O.unet_predictor(sd, a, b), steps, noises)
    print("unet64 %d-step DDPM sample: rel_l2 vs CPU oracle = %.3e" % (steps, rel_l2(y, r)))
