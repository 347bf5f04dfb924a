"""CPU study: relative error of one UNet forward when conv operands are rounded
to TF32 (1 MMA) or split into bf16 hi+lo (3 MMAs), vs strict fp32.
Decides the tensor-core operand format ('Precision')."""
import sys, torch
sys.path[:0] = [".", "tests"]
from oracle import hotpath as O
from helpers import model_sd
from vq_voice_swap_b200 import synth
import torch.nn.functional as F

def tf32(x):
    i = x.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF   # round-to-nearest (ties away) to 10 mantissa bits
    return i.view(torch.float32)

def bf16_split(x):
    hi = x.bfloat16().float()
    lo = (x - hi).bfloat16().float()
    return hi, lo

mode = None
orig_conv = O.conv
def conv(x, sd, key, dilation=1):
    w = sd[key + ".weight"]; b = sd[key + ".bias"]
    pad = dilation * (w.shape[-1] // 2)
    if mode is None or x.shape[1] < 16:
        return F.conv1d(x, w, b, padding=pad, dilation=dilation)
    if mode == "tf32":
        return F.conv1d(tf32(x.contiguous()), tf32(w.contiguous()), b, padding=pad, dilation=dilation)
    if mode == "bf16":
        return F.conv1d(x.bfloat16().float(), w.bfloat16().float(), b, padding=pad, dilation=dilation)
    xh, xl = bf16_split(x); wh, wl = bf16_split(w)
    y = F.conv1d(xh, wh, b, padding=pad, dilation=dilation)
    y = y + F.conv1d(xl, wh, None, padding=pad, dilation=dilation) + F.conv1d(xh, wl, None, padding=pad, dilation=dilation)
    return y
O.conv = conv

torch.manual_seed(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sd = model_sd("diffusion_unet32", "study32")
x = synth.normal("study/x", (1, 1, T))
ts = torch.tensor([0.5])
with torch.no_grad():
    ref = O.unet_predictor(sd, x, ts)
    for m in ["tf32", "bf16x3", "bf16"]:
        mode = m
        y = O.unet_predictor(sd, x, ts)
        print(m, "rel_l2 = %.3e" % float((y - ref).norm() / ref.norm()), "max_abs = %.3e" % float((y - ref).abs().max()), "ref rms %.3f" % float(ref.pow(2).mean().sqrt()), flush=True)
