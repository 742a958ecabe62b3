#!/usr/bin/env python3
"""Per-INSTRUCTION reading of an `ncu --page source --csv --print-source cuda,sass` export (tools/gpu_prof.sh leaves it as
gpurun_out/<tag>_src.csv): calls that are actually executed (slow-path routines show up here), the share of warp
instructions by active lanes, contiguous low-lane regions, and the instructions that own the most stall samples.
    tools/ncu_lowlane.py gpurun_out/r2g_k_walk_src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur, seen = None, {}
for row in rows:
    if len(row) >= 2 and row[0] == 'File Path':
        cur = row[1]
        continue
    if len(row) > 8 and row[2].startswith('0x'):
        a = int(row[2], 16)
        if a not in seen:   # (file, address, sass, stall samples, warp instructions, thread instructions)
            seen[a] = ((cur or '?').split('/')[-1], a, row[3].strip(), int(row[4] or 0), int(row[7] or 0), int(row[8] or 0))
ins = sorted(seen.values(), key=lambda s: s[1])
base = ins[0][1]
tot = sum(s[4] for s in ins)
stot = sum(s[3] for s in ins)
print("warp instructions %.4g   static instructions %d   stall samples %d" % (tot, len(ins), stot))
print("executed calls:")
for s in ins:
    if s[2].startswith('CALL') and s[4] > 0:
        print("  %6x  %10d calls  %4.1f lanes  %s" % (s[1] - base, s[4], s[5] / s[4], s[0]))
bins = collections.Counter()
for s in ins:
    if s[4]:
        bins[min(32, int((s[5] / s[4]) // 4 * 4))] += s[4]
print("share of warp instructions by active lanes (bin = lower edge):", {k: round(100 * v / tot, 1) for k, v in sorted(bins.items())})
reg, r = [], None
for s in ins:
    low = s[4] > 0 and s[5] / s[4] < 12
    if low:
        if r is None:
            r = [s[1], s[1], 0, 0, 0, collections.Counter()]
        r[1] = s[1]; r[2] += s[4]; r[3] += s[5]; r[4] += s[3]; r[5][s[0]] += s[4]
    elif r and s[4] > 0:
        reg.append(r); r = None
if r:
    reg.append(r)
reg.sort(key=lambda x: -x[2])
print("regions below 12 lanes (%.1f %% of all warp instructions):" % (100 * sum(x[2] for x in reg) / tot))
for x in reg[:8]:
    print("  %6x-%6x  %5.2f %% of instructions  %5.2f %% of stalls  %4.1f lanes  %4d instructions  %s" %
          (x[0] - base, x[1] - base, 100 * x[2] / tot, 100 * x[4] / stot, x[3] / x[2], (x[1] - x[0]) // 16 + 1, x[5].most_common(2)))
print("instructions with the most stall samples:")
for s in sorted(ins, key=lambda s: -s[3])[:14]:
    print("  %6x  %5.2f %%  %10d executions  %4.1f lanes  %-16s %s" % (s[1] - base, 100 * s[3] / stot, s[4], s[5] / max(s[4], 1), s[0], s[2][:60]))
