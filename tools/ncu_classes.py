"""Run one ResBlock per conv shape class of the unet64 step at the benchmark batch, so that
`ncu --set full --profile-from-start off -k regex:conv_umma` captures exactly two launches (conv1, conv2) per class.

usage: ncu ... python tools/ncu_classes.py [--batch 64] [--only name,name]
Each class prints `class <name> launches <first>..<last>` so the report's launch indices map back to shapes."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vq_voice_swap_b200 import synth
from vq_voice_swap_b200.unet import ResBlock

# name: (c_in, c_out, t_in, scale, dilation)
CLASSES = {
    "c64_l0":        (64, 64, 64000, 1.0, 2),     # conv1 64->64 resident, conv2 identity skip
    "c128to64_l0":   (128, 64, 64000, 1.0, 2),    # concat-fed up block: conv1 128->64, conv2 64->64 + 1x1 skip(128)
    "c64_down_l0":   (64, 64, 64000, 0.5, 2),
    "c64_up_l1":     (64, 64, 32000, 2.0, 2),
    "c192to64_l1":   (192, 64, 32000, 1.0, 2),
    "c128_l2":       (128, 128, 16000, 1.0, 2),   # conv1 128->128 streamed weights, conv2 identity skip
    "c256to128_l2":  (256, 128, 16000, 1.0, 2),   # conv1 256->128, conv2 128->128 + 1x1 skip(256)
    "c128_up_l2":    (128, 128, 16000, 2.0, 2),   # 16000 -> 32000
    "c128_down_l2":  (128, 128, 16000, 0.5, 2),
    "c64to128_l2":   (64, 128, 16000, 1.0, 2),
    "c256_l5":       (256, 256, 2000, 1.0, 2),
    "c512to256_l5":  (512, 256, 2000, 1.0, 2),
    "c512_l8":       (512, 512, 250, 1.0, 2),
    "c1024to512_l8": (1024, 512, 250, 1.0, 2),
    "c512_mid_d16":  (512, 512, 250, 1.0, 16),
}

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--only", default="")
a = ap.parse_args()
names = [n for n in a.only.split(",") if n] or list(CLASSES)
launch = 0
for name in names:
    cin, cout, t, scale, dil = CLASSES[name]
    # a standalone block has no network around it to derive the operand format from: apply the unet64 rule (engine.conv_precision)
    os.environ["VQVS_PREC"] = "f16" if cout >= 256 else "bf16x3"
    blk = ResBlock(cin, 256, cout, scale_factor=scale, dilation=dil)
    synth.load_synth(blk, "runblock")
    blk = blk.cuda()
    x = torch.randn(a.batch, cin, t, device="cuda")
    emb = torch.randn(a.batch, 256, device="cuda")
    blk(x, emb)  # plan build + weight packing (not captured)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = blk(x, emb)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("class %s cin=%d cout=%d t=%d scale=%g dil=%d launches %d..%d std=%.4f" % (name, cin, cout, t, scale, dil, launch, launch + 1, float(y.std())), flush=True)
    launch += 2
    del blk, x, emb, y
    torch.cuda.empty_cache()
