"""Run one ResBlock (for ncu captures of the conv kernels at a chosen shape)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vq_voice_swap_b200 import synth
from vq_voice_swap_b200.unet import ResBlock

ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=64)
ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--t", type=int, default=64000)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--dilation", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
blk = ResBlock(a.cin, 256, a.cout, scale_factor=a.scale, dilation=a.dilation)
synth.load_synth(blk, "runblock")
blk = blk.cuda()
x = torch.randn(a.batch, a.cin, a.t, device="cuda")
emb = torch.randn(a.batch, 256, device="cuda")
for _ in range(a.reps):
    y = blk(x, emb)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.std()))
