#!/bin/bash
out=gpurun_out/x5; mkdir -p $out
D=$PWD/vq_voice_swap_b200
timeout 300 python tools/op_profile.py > $out/op_def.txt 2>&1; echo "def: $(sed -n 2p $out/op_def.txt)"
VQVS_NO_STACK=1 timeout 300 python tools/op_profile.py > $out/op_nostack.txt 2>&1; echo "nostack: $(sed -n 2p $out/op_nostack.txt)"
VQVS_NO_STACK=1 VQVS_LIB=$D/libvqvs_coll.so timeout 300 python tools/op_profile.py > $out/op_nostack_coll.txt 2>&1; echo "nostack+coll: $(sed -n 2p $out/op_nostack_coll.txt)"
timeout 300 python tools/op_profile.py > $out/op_def2.txt 2>&1; echo "def: $(sed -n 2p $out/op_def2.txt)"
for pr in default fast; do timeout 300 python bench.py --precision $pr --steps 2 --warmup 3 --no-eager --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['precision'], d['value'], d['clocks'], d['roofline']['whole_path_frac'])"; done
