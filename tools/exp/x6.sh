#!/bin/bash
D=$PWD/vq_voice_swap_b200
run() { "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks']['sm_mhz'], d['roofline']['whole_path_frac'])"; }
echo "def: $(run env)"
echo "nostack: $(run env VQVS_NO_STACK=1)"
echo "nostack+coll: $(run env VQVS_NO_STACK=1 VQVS_LIB=$D/libvqvs_coll.so)"
echo "coll: $(run env VQVS_LIB=$D/libvqvs_coll.so)"
echo "def: $(run env)"
