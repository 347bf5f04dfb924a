"""Kernel-boundary cost of conv_umma_kernel: one conv launched 50x back to back (no events in between) against the cycles
CTA 0 spends inside the kernel (role profiler build, VQVS_LIB=libvqvs_prof.so).  usage: gap_probe.py cin,cout,t,batch ..."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vq_voice_swap_b200 import synth, lib as L
from vq_voice_swap_b200.unet import ResBlock

lib = L.load()
for spec in sys.argv[1:]:
    cin, cout, t, batch = [int(v) for v in spec.split(",")]
    blk = ResBlock(cin, 256, cout, scale_factor=1.0, dilation=2)
    synth.load_synth(blk, "runblock"); blk = blk.cuda()
    x = torch.randn(batch, cin, t, device="cuda"); emb = torch.randn(batch, 256, device="cuda")
    blk(x, emb); blk(x, emb)
    plan = next(iter(blk._plans.items.values()))
    conv = [(k, d) for k, d in plan.descs if k == L.OP_CONV_UMMA][0]
    reps = 50
    ops = (L.Op * reps)()
    for i in range(reps):
        ops[i].kind = conv[0]; ops[i].desc = C.addressof(conv[1])
    conv[1].reserved_ |= 512
    L.check(lib.vqvs_run(ops, reps, L.stream_ptr())); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.check(lib.vqvs_run(ops, reps, L.stream_ptr())); e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) * 1e3 / reps
    pb = (C.c_uint64 * 32)(); L.check(lib.vqvs_debug_prof(pb))
    clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
    print("%s: %.2f us per launch back to back; CTA 0 inside the kernel %d clocks (%.2f us at %d MHz), prologue+grid wait %d clocks" %
          (spec, per, pb[5], pb[5] / clk, clk, pb[6]), flush=True)
    if hasattr(lib, "vqvs_debug_cta"):  # per-CTA timeline of the last launch (profiling build)
        import numpy as np
        def timeline():
            buf = (C.c_uint64 * 640)(); lib.vqvs_debug_cta(buf)
            a = np.array(list(buf), dtype=np.int64).reshape(160, 4)
            return a[a[:, 3] > 0]
        one = (L.Op * 1)(); one[0].kind = conv[0]; one[0].desc = C.addressof(conv[1])
        runs = []
        for _ in range(3):
            L.check(lib.vqvs_run(one, 1, L.stream_ptr())); torch.cuda.synchronize()
            runs.append(timeline())
        for a in runs:
            dur = (a[:, 2] - a[:, 1]) / 1e3
            end = (a[:, 2] - a[:, 2].min()) / 1e3
            print("   CTAs %d: busy us min %.1f median %.1f max %.1f; end-time spread %.1f us (median end %.1f before the last); tiles %d..%d" %
                  (len(a), dur.min(), np.median(dur), dur.max(), end.max(), end.max() - np.median(end), a[:, 3].min(), a[:, 3].max()))
        # is slowness a property of the SM?  correlation of per-SM busy time between launches
        d0 = {int(r[0]): (r[2] - r[1]) / r[3] for r in runs[0]}; d1 = {int(r[0]): (r[2] - r[1]) / r[3] for r in runs[1]}
        k = sorted(set(d0) & set(d1))
        print("   per-SM time per tile, launch 1 vs 2: correlation %.2f over %d SMs; slowest/fastest SM %.3f" %
              (np.corrcoef([d0[i] for i in k], [d1[i] for i in k])[0, 1], len(k), max(d0.values()) / min(d0.values())))
    conv[1].reserved_ &= ~512
    del blk, x, plan
