import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from vq_voice_swap_b200 import synth, lib as L
from vq_voice_swap_b200.unet import ResBlock
ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--t", type=int, default=64000); ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--dilation", type=int, default=2)
a = ap.parse_args()
blk = ResBlock(a.cin, 256, a.cout, scale_factor=a.scale, dilation=a.dilation)
synth.load_synth(blk, "runblock"); blk = blk.cuda()
x = torch.randn(a.batch, a.cin, a.t, device="cuda"); emb = torch.randn(a.batch, 256, device="cuda")
blk(x, emb); blk(x, emb)
plan = next(iter(blk._plans.items.values()))
n = len(plan.descs); buf = (C.c_float * n)(); acc = [0.0] * n
for _ in range(5):
    L.check(L.load().vqvs_run_timed(plan.ops, n, L.stream_ptr(), buf))
    for i in range(n): acc[i] += buf[i] / 5
print(" ".join("%s=%.3fms" % ({2: "umma", 1: "simt", 3: "gn"}[k], ms) for (k, _), ms in zip(plan.descs, acc)))

if int(os.environ.get("VQVS_DEBUG_FLAGS", "0")) & 512:
    buf = (C.c_uint64 * 32)()
    L.check(L.load().vqvs_debug_prof(buf))
    names = ["xf.wait_ab", "xf.wait_raw", "xf.work", "xf.loop", "tma.wait_empty", "tma.issue", "tma.-", "tma.-", "mma.wait_a", "mma.issue", "mma.wait_acc", "mma.loop", "epi.wait_full", "epi.stats+next", "epi.tmem_ld", "epi.store"]
    print(" | ".join("%s=%d" % (n, buf[i]) for i, n in enumerate(names) if buf[i]))
