"""Per-kernel timing of one ResBlock's conv kernels under ablation flags, plus the role profiler
(libvqvs built with -DVQVS_PROF, flag 512).  usage: prof_roles.py cin,cout,t,scale,dil[,batch] ..."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vq_voice_swap_b200 import synth, lib as L
from vq_voice_swap_b200.unet import ResBlock

FLAGS = [int(x) for x in os.environ.get("ABLATE", "0,1,8,16,64,80,4").split(",")]
PROF = os.environ.get("PROF", "0") == "1"
NAMES = ["xf.wait_ab", "xf.wait_raw", "xf.work", "xf.loop", "mma.wait_b", "cta.total", "cta.prologue", "tma.3", "mma.wait_ab", "mma.issue",
         "mma.wait_acc", "mma.loop", "epi.wait_full", "epi.stats+next", "epi.tmem_ld", "epi.store"]
lib = L.load()
for spec in sys.argv[1:]:
    f = spec.split(",")
    cin, cout, t = int(f[0]), int(f[1]), int(f[2])
    scale, dil = float(f[3]), int(f[4])
    batch = int(f[5]) if len(f) > 5 else 16
    blk = ResBlock(cin, 256, cout, scale_factor=scale, dilation=dil)
    synth.load_synth(blk, "runblock"); blk = blk.cuda()
    x = torch.randn(batch, cin, t, device="cuda"); emb = torch.randn(batch, 256, device="cuda")
    blk(x, emb); blk(x, emb)
    plan = next(iter(blk._plans.items.values()))
    for i, (kind, d) in enumerate(plan.descs):
        if kind != L.OP_CONV_UMMA:
            continue
        one = (L.Op * 1)(); one[0].kind = kind; one[0].desc = C.addressof(d)
        alg = 4.0 * d.batch * ((d.c_a + d.c_b) * d.t_in + d.c_out * d.t_out + ((d.s_a + d.s_b) * d.t_skip if d.skip_mode else 0))
        line = []; prof = ''
        keep = d.reserved_ & ~0x3FF  # VQVS_CONV_PAIR_STATS, statistics granularity and operand format set by the engine
        for fl in FLAGS:
            d.reserved_ = fl | keep | (512 if PROF else 0)
            ms = 0.0; buf = (C.c_float * 1)()
            for r in range(4):
                L.check(lib.vqvs_run_timed(one, 1, L.stream_ptr(), buf))
                if r: ms += buf[0] / 3
            line.append("f%d=%.3fms(%.0fGB/s)" % (fl, ms, alg / ms / 1e6))
            if PROF:
                pb = (C.c_uint64 * 32)(); L.check(lib.vqvs_debug_prof(pb))
                prof += "\n    prof f%d: " % fl + " | ".join("%s=%d" % (n, pb[j]) for j, n in enumerate(NAMES) if pb[j])
        d.reserved_ = keep
        print("%s op%d cin=%d cout=%d tin=%d tout=%d k=%d d=%d skip=%d(%d): %s" % (spec, i, d.c_a + d.c_b, d.c_out, d.t_in, d.t_out,
              d.ksize, d.dilation, d.skip_mode, d.s_a + d.s_b, "  ".join(line)))
        if PROF: print(prof)
    del blk, x, plan
    torch.cuda.empty_cache()
