#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares from
   ncu -i rep --page source --print-source sass,cuda --csv --kernel-name regex:K > x.csv
   usage: tools/ncu_lines.py x.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
fname = ''
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if len(r) > 3 and r[0] == 'Line No':
        hdr = r
        ii = [i for i, h in enumerate(hdr) if h == 'Instructions Executed'][0]
        si = [i for i, h in enumerate(hdr) if h == '# Samples'][0]
        continue
    if hdr and len(r) == len(hdr) and r[2] == '-':  # source-line aggregate rows
        try:
            ie = float(r[ii] or 0)
            ss = float(r[si] or 0)
        except ValueError:
            continue
        k = (fname, r[0], r[1].strip()[:95])
        a = agg.setdefault(k, [0.0, 0.0])
        a[0] += ie
        a[1] += ss
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print(f'total warp-inst {tot:.3e}  samples {tots:.0f}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{a[0] / tot * 100:5.1f}% inst {a[1] / tots * 100:5.1f}% smp  {k[0]}:{k[1]:>4}  {k[2]}')
