"""Forward error (vs the fp32 CPU oracle, unet64, T = 64000) and conv time per UNet step (batch 64) of operand-format policies:
VQVS_F16_FROM (fp16 single products from this multiple of base_channels) x VQVS_BF16X3_FIRST / _LAST (blocks at the ends of the
network that keep the bf16 hi/lo split).  usage: python tools/precision_policy_study.py "from,first,last" ..."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from helpers import rel_l2
from oracle import hotpath as O
from vq_voice_swap_b200 import synth
from vq_voice_swap_b200.diffusion_model import DiffusionModel
import bench

os.environ["VQVS_BACKEND"] = "umma"
m = DiffusionModel("unet", 64)
sd = synth.synth_state_dict(synth.shapes_of(m), tag="full64"); m.load_state_dict(sd); m = m.cuda().eval()
x = synth.normal("full64/x", (1, 1, 64000)); ts = torch.tensor([0.62])
ref = O.unet_predictor(sd, x, ts)
xb = torch.randn(64, 1, 64000, device="cuda")
for spec in sys.argv[1:]:
    f, first, last, *blocks = spec.split(",")
    os.environ["VQVS_F16_FROM"], os.environ["VQVS_BF16X3_FIRST"], os.environ["VQVS_BF16X3_LAST"] = f, first, last
    os.environ["VQVS_F16_BLOCKS"] = ",".join(blocks)
    m.predictor._plans.clear()
    got = m.predictor(x.cuda(), ts.cuda()).cpu()
    err = rel_l2(got, ref)
    plan, per_op, by_kind, umma = bench.profile_kernels(m, xb)
    print("from=%s first=%s last=%s blocks=%s: forward rel_l2 = %.3e, conv %.3f ms per UNet step" % (f, first, last, ",".join(blocks), err, by_kind.get(2, 0.0)), flush=True)
    torch.cuda.empty_cache()
