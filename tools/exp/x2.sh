#!/bin/bash
out=gpurun_out/x2; mkdir -p $out
for f in 4 2 1; do
  echo "== F16_FROM=$f"
  VQVS_F16_FROM=$f timeout 300 python tools/op_profile.py > $out/op_f$f.txt 2>&1; sed -n 1,2p $out/op_f$f.txt
  VQVS_F16_FROM=$f timeout 900 python tools/measure_parity.py 4 50 2>&1 | tail -3
done 2>&1 | tee $out/log.txt
