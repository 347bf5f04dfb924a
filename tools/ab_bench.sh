#!/bin/bash
# Sustained A/B of library builds through bench.py (the step runs under the power cap: short per-layer timings mislead).
# usage: tools/ab_bench.sh [name=path ...]   (default: prev=libvqvs_prev.so against the shipped libvqvs.so)
D=$PWD/vq_voice_swap_b200
run() { "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks']['sm_mhz'], d['roofline']['whole_path_frac'])"; }
[ $# -eq 0 ] && set -- prev=$D/libvqvs_prev.so
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -1
for round in 1 2; do
  for v in "$@"; do echo "${v%%=*}: $(run env VQVS_LIB=${v#*=})"; done
  echo "new: $(run env)"
done
