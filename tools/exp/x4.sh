#!/bin/bash
out=gpurun_out/x4; mkdir -p $out
timeout -s KILL 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_umma -o $out/cls -f \
    python tools/ncu_classes.py --only c128to64_l0,c128_l2 > $out/classes.log 2>&1
grep '^class' $out/classes.log
ncu -i $out/cls.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
for i in 0 2; do
  ncu -i $out/cls.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | python tools/ncu_source_filter.py | gzip > $out/source_$i.csv.gz
done
rm -f $out/cls.ncu-rep; ls -la $out
