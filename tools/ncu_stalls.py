#!/usr/bin/env python3
"""Stall-reason totals of a kernel and the lines that own each reason, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`:   ncu_stalls.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cur, hdr = None, None
tot = collections.Counter()
by = collections.defaultdict(collections.Counter)
smem = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    for k in hdr:
        if k.startswith('stall_') and '(Not Issued)' not in k:
            try:
                v = float(d[k].replace(',', ''))
            except ValueError:
                v = 0.0
            if v:
                tot[k] += v
                by[k][(cur, r[0], r[1].strip()[:70])] += v
    for k in ('L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal'):
        try:
            smem[k] += float(d.get(k, '0').replace(',', ''))
        except ValueError:
            pass
s = sum(tot.values())
print("stall samples by reason (all samples):")
for k, v in tot.most_common():
    print("  %-24s %6.2f%%" % (k, 100 * v / s))
    for (f, ln, src), w in by[k].most_common(top if v / s > 0.04 else 0):
        print("      %5.2f%%  %-16s %4s  %s" % (100 * w / s, f, ln, src))
print("shared-memory wavefronts: %.3e, ideal %.3e" % (smem['L1 Wavefronts Shared'], smem['L1 Wavefronts Shared Ideal']))
