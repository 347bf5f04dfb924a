"""Per-kernel device times of the guidance model's forward and input-gradient launch programs (BASELINE configs[4]:
Classifier bc32, batch 32, T = 64000).  usage: python tools/guidance_profile.py [batch]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vq_voice_swap_b200 import guidance, lib as L

dev = torch.device("cuda:0")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
clf = bench.build_classifier(dev) if hasattr(bench, "build_classifier") else None
if clf is None:
    from vq_voice_swap_b200.classifier import Classifier
    from vq_voice_swap_b200 import synth
    clf = Classifier(num_labels=100, base_channels=32)
    synth.load_synth(clf, "bench/classifier"); clf = clf.to(dev).eval()
x = torch.randn(batch, 1, 64000, device=dev); ts = torch.full((batch,), 0.5, device=dev)
plans = guidance.classifier_plans(clf, x)
plans.forward(x, ts); plans.backward(torch.ones(batch, clf.num_labels, device=dev), plans.generation)
lib = L.load()
for name, plan in (("forward", plans.fwd), ("backward", plans.bwd)):
    n = len(plan.descs); buf = (C.c_float * n)(); acc = [0.0] * n
    for _ in range(3):
        L.check(lib.vqvs_run_timed(plan.ops, n, L.stream_ptr(), buf))
        for i in range(n): acc[i] += buf[i] / 3
    by = {}
    for (k, d), ms in zip(plan.descs, acc):
        e = by.setdefault(L.OP_NAMES.get(k, str(k)), [0, 0.0]); e[0] += 1; e[1] += ms
    print("%s: %d launches, %.3f ms" % (name, n, sum(acc)))
    for k, (cnt, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("   %-16s %4d launches %8.3f ms" % (k, cnt, ms))
    if name == "backward":
        rows = []
        for (k, d), ms in zip(plan.descs, acc):
            if k == L.OP_CONV_UMMA:
                rows.append((ms, "conv^T cin=%d cout=%d t=%d->%d" % (d.c_a + d.c_b, d.c_out, d.t_in, d.t_out)))
            elif k == L.OP_GELU_BWD:
                rows.append((ms, "gelu_bwd c=%d t=%d up=%d  (%.0f GB/s)" % (d.c, d.t, d.up, 12.0 * d.batch * d.c * d.t / ms / 1e6)))
            elif k == L.OP_AFFINE3:
                rows.append((ms, "affine3 c=%d t=%d add_mode=%d add2=%d  (%.0f GB/s)" % (d.c, d.t, d.add_mode, 1 if d.add2 else 0,
                             (12.0 + (4 if d.add_mode else 0) + (4 if d.add2 else 0)) * d.batch * d.c * d.t / ms / 1e6)))
        for ms, what in sorted(rows, reverse=True)[:24]:
            print("      %.3f ms  %s" % (ms, what))
