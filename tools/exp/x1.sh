#!/bin/bash
# experiment round: GELU degree A/B, cooperative finalize, role profile of the production kinds
out=gpurun_out/x1; mkdir -p $out
D=$PWD/vq_voice_swap_b200
timeout 600 python -m pytest tests -q -m gpu -x --timeout 200 2>&1 | tail -2 > $out/pytest.txt; cat $out/pytest.txt
for v in base g0 def g6 base def; do
  lib=$D/libvqvs_$v.so; [ $v = def ] && lib=$D/libvqvs.so
  VQVS_LIB=$lib timeout 300 python tools/op_profile.py > $out/op_$v.txt 2>&1; echo "$v: $(sed -n 2p $out/op_$v.txt)"
done
for v in def g6 g0; do
  lib=$D/libvqvs_$v.so; [ $v = def ] && lib=$D/libvqvs.so
  echo "parity $v"; VQVS_LIB=$lib timeout 600 python tools/measure_parity.py 4 2>&1 | tail -2
done > $out/parity.txt 2>&1; cat $out/parity.txt
SH="64,64,64000,1,2,64 128,64,64000,1,2,64 128,128,16000,1,2,64 256,128,16000,1,2,64"
VQVS_LIB=$D/libvqvs_prof.so PROF=1 ABLATE=0 timeout 300 python tools/prof_roles.py $SH > $out/roles.txt 2>&1
SH2="256,256,2000,1,2,64 512,256,2000,1,2,64 512,512,500,1,2,64 512,512,250,1,2,64 1024,512,250,1,2,64"
VQVS_PREC=f16 VQVS_LIB=$D/libvqvs_prof.so PROF=1 ABLATE=0 timeout 300 python tools/prof_roles.py $SH2 >> $out/roles.txt 2>&1
cat $out/roles.txt
