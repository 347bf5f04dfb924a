"""stdin: `ncu --page source --csv` of one launch; stdout: the columns used by tools/ncu_by_role.py, one row per SASS instruction."""
import csv, sys
rows = list(csv.reader(sys.stdin))
w = csv.writer(sys.stdout)
if len(rows) > 2:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    keep = [k for k in ['Address', 'Source', '# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier',
            'stall_math', 'stall_mio', 'stall_no_inst', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_branch_resolving',
            'stall_lg', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal'] if k in ix]
    w.writerow([rows[0][1] if len(rows[0]) > 1 else ''])
    w.writerow(keep)
    seen = set()
    for r in rows[2:]:
        if len(r) < len(hdr) or r[ix['Address']] in seen: continue
        seen.add(r[ix['Address']])
        w.writerow([r[ix[k]] for k in keep])
