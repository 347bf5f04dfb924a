#!/bin/bash
# Per-shape-class ncu capture of conv_umma_kernel (two launches per class, tools/ncu_classes.py).
# usage: tools/ncu_classes.sh <outdir> [--only a,b,c]
out=${1:-gpurun_out/ncu_classes}; shift
mkdir -p $out
timeout -s KILL 900 ncu --set full --clock-control none --profile-from-start off -k regex:conv_umma -o $out/classes -f \
    python tools/ncu_classes.py "$@" > $out/classes.log 2>&1
grep '^class' $out/classes.log
ncu -i $out/classes.ncu-rep --page raw --csv > $out/classes_raw.csv 2>/dev/null
# SASS-level source page per launch, keeping only instructions that were sampled or executed a lot
n=$(grep -c '^class' $out/classes.log); n=$((2 * n))
for i in $(seq 0 $((n - 1))); do
  ncu -i $out/classes.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
w = csv.writer(sys.stdout)
if len(rows) > 2:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    keep = ['Address', 'Source', '# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier',
            'stall_math', 'stall_mio', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_branch_resolving',
            'stall_lg', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal', 'L2 Theoretical Sectors Global']
    keep = [k for k in keep if k in ix]
    w.writerow([rows[0][1] if len(rows[0]) > 1 else ''])
    w.writerow(keep)
    seen = set()
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ix['Address']] in seen: continue
        seen.add(r[ix['Address']])
        w.writerow([r[ix[k]] for k in keep])
" | gzip > $out/source_$i.csv.gz
done
sz=$(stat -c %s $out/classes.ncu-rep)
if [ $sz -gt 30000000 ]; then rm $out/classes.ncu-rep; echo "report was $sz bytes: removed (csv pages kept)"; fi
ls -la $out
