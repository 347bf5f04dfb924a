#!/bin/bash
out=gpurun_out/x9; mkdir -p $out
run() { name=$1; shift; env "$@" timeout 300 python tools/op_profile.py > $out/op_$name.txt 2>&1; echo "$name: $(sed -n 2p $out/op_$name.txt)"; }
run def A=1
run kbs1 VQVS_FORCE_KBS=1
run kbs2 VQVS_FORCE_KBS=2
run kbs4 VQVS_FORCE_KBS=4
run res150 VQVS_RESIDENT_KB=150
run res60 VQVS_RESIDENT_KB=60
run def2 A=1
