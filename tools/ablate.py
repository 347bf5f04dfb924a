"""Time one ResBlock's two conv kernels under ablation flags (profiling aid)."""
import os, sys, subprocess
flags = [int(x) for x in os.environ.get('ABLATE', '0,1,16,31,63,95,127').split(',')]
for f in flags:
    env = dict(os.environ, VQVS_DEBUG_FLAGS=str(f))
    out = subprocess.run([sys.executable, "tools/time_block.py"] + sys.argv[1:], env=env, capture_output=True, text=True)
    print("flags=%2d  %s" % (f, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]))
