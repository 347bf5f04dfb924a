#!/bin/bash
out=gpurun_out/r2more; mkdir -p $out
timeout -s KILL 1200 ncu --set full --clock-control none --profile-from-start off -k regex:conv_umma -o $out/classes -f \
    python tools/ncu_classes.py --only c64_up_l1,c192to64_l1,c128_down_l2,c64to128_l2,c512to256_l5,c1024to512_l8,c512_mid_d16 > $out/classes.log 2>&1
grep '^class' $out/classes.log
ncu -i $out/classes.ncu-rep --page raw --csv > $out/classes_raw.csv 2>/dev/null
rm -f $out/classes.ncu-rep; ls -la $out
