#!/bin/bash
# quick GPU iteration: parity tests + per-kernel timings of the main ResBlock shapes.  usage: tools/gpu_quick.sh <tag> [ABLATE list]
tag=${1:-q}; out=gpurun_out/$tag; mkdir -p $out
SH="64,64,64000,1,2 128,64,64000,1,2 64,64,32000,2,2 64,64,64000,0.5,2 128,128,16000,1,2 256,128,16000,1,2 256,256,2000,1,2 512,512,500,1,2,64 1024,512,500,1,2,64"
timeout 400 python -m pytest tests -q -m gpu -x --timeout 120 > $out/pytest.txt 2>&1; tail -3 $out/pytest.txt
ABLATE=${2:-0} timeout 300 python tools/prof_roles.py $SH > $out/ablate.txt 2>&1; cat $out/ablate.txt
