#!/bin/bash
D=$PWD/vq_voice_swap_b200
run() { "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['clocks']['sm_mhz'], d['roofline']['whole_path_frac'])"; }
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -1
echo "prev: $(run env VQVS_LIB=$D/libvqvs_prev.so)"
echo "new: $(run env)"
echo "prev: $(run env VQVS_LIB=$D/libvqvs_prev.so)"
echo "new: $(run env)"
