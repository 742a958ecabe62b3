#!/usr/bin/env python3
"""Per-source-line instruction counts and lane efficiency from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] in ('Function Name',) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    def num(k):
        try:
            return float(d.get(k, '0').replace(',', ''))
        except ValueError:
            return 0.0
    lines.append((num('Instructions Executed'), num('Thread Instructions Executed'), num('Warp Stall Sampling (All Samples)'),
                  cur_file, r[0], r[1].strip()[:90]))
tot_i = sum(l[0] for l in lines)
tot_t = sum(l[1] for l in lines)
tot_s = sum(l[2] for l in lines)
print("total warp-instr %.3e  thread-instr %.3e  lanes/instr %.2f  stall samples %d" % (tot_i, tot_t, tot_t / max(tot_i, 1), tot_s))
print("%7s %7s %6s  %-18s %5s  %s" % ("instr%", "stall%", "lanes", "file", "line", "source"))
for l in sorted(lines, reverse=True)[:top]:
    print("%6.2f%% %6.2f%% %6.1f  %-18s %5s  %s" % (100 * l[0] / tot_i, 100 * l[2] / max(tot_s, 1), l[1] / max(l[0], 1), l[3], l[4], l[5]))
