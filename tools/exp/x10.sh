#!/bin/bash
out=gpurun_out/x10; mkdir -p $out
D=$PWD/vq_voice_swap_b200
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 2>&1 | tail -2
for v in prev def prev def; do
  lib=$D/libvqvs_$v.so; [ $v = def ] && lib=$D/libvqvs.so
  VQVS_LIB=$lib timeout 300 python tools/op_profile.py > $out/op_$v.txt 2>&1; echo "$v: $(sed -n 2p $out/op_$v.txt)"
done
