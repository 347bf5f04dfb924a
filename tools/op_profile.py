"""Per-op device times of one UNet step (CUDA events between launches): which layers are slow."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vq_voice_swap_b200 import lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--bc", type=int, default=64)
ap.add_argument("--t", type=int, default=64000)
a = ap.parse_args()
dev = torch.device("cuda:0")
model = bench.build_model(dev, a.bc)
x = torch.randn(a.batch, 1, a.t, device=dev)
plan, per_op, by_kind, umma = bench.profile_kernels(model, x)
alg = bench.conv_algorithmic_bytes(plan)
names = {1: "conv_simt", 2: "conv_umma", 3: "gn_finalize", 4: "conv_in", 5: "conv_out", 6: "time_embed", 7: "film", 8: "memset", 9: "ddpm_finish"}
print("total ms per UNet step: %.2f" % sum(per_op))
for k, v in sorted(by_kind.items()):
    print("  %-12s %8.3f ms" % (names[k], v))
agg = {}
for (kind, d), ms, b in zip(plan.descs, per_op, alg):
    if kind in (1, 2):
        key = (d.c_a + d.c_b, d.c_out, d.t_in, d.t_out, d.ksize, d.dilation, d.skip_mode, d.s_a + d.s_b)
        e = agg.setdefault(key, [0, 0.0, b, 0.0])
        e[0] += 1; e[1] += ms
        flops = 2.0 * d.batch * d.t_out * d.c_out * ((d.c_a + d.c_b) * d.ksize + (d.s_a + d.s_b if d.skip_mode == 2 else 0))
        e[3] = flops
print("%5s %5s %6s %6s k d skip cs |  n   ms/launch   GB/s(alg)  TFLOP/s(fp32-equiv)  share" % ("cin", "cout", "tin", "tout"))
tot = sum(per_op)
for key, (n, ms, b, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    per = ms / n
    print("%5d %5d %6d %6d %d %2d %d %4d | %2d  %8.3f  %9.1f  %9.1f   %5.1f%%" % (*key, n, per, b / per / 1e6, fl / per / 1e9, 100 * ms / tot))
