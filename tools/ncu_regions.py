#!/usr/bin/env python3
"""Instruction and stall-sample shares per source FUNCTION of a kernel, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv`:   ncu_regions.py src.csv
Lines are attributed to the function whose definition encloses them in the repository's sources."""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = {f: os.path.join(ROOT, "mc_old_b200/csrc", f) for f in ("mcb_kernels.cu", "mcb_walk.cu", "mcb_events.cuh", "mcb_device.cuh",
                                                              "mcb_physics.h", "mcb_tables.h")}
DEF = re.compile(r'^(?:static\s+)?(?:template\s*<[^>]*>\s*)?(?:__global__|__device__|MCB_HD|MCB_THD)[^;(]*?\b(\w+)\s*\(')


def function_map(path):
    """line number -> name of the last function definition that started at or before it"""
    out, cur = {}, "?"
    pending = ""
    for i, line in enumerate(open(path), 1):
        text = line.strip()
        if text.startswith("template") and "(" not in text:
            pending = text + " "
            out[i] = cur
            continue
        m = DEF.match(re.sub(r'__launch_bounds__\([^)]*\)', '', pending + text))
        if m is None and "__global__" in (pending + text) and "(" not in re.sub(r'__launch_bounds__\([^)]*\)', '', pending + text):
            pending = pending + text + " "
            out[i] = cur
            continue
        pending = ""
        if m:
            cur = m.group(1)
        out[i] = cur
    return out


MAPS = {f: function_map(p) for f, p in SRC.items() if os.path.exists(p)}
rows = list(csv.reader(open(sys.argv[1])))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
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

    def f(k):
        try:
            return float(d.get(k, '0').replace(',', ''))
        except ValueError:
            return 0.0
    name = MAPS.get(cur, {}).get(int(r[0]), cur)
    a = agg[name if cur in MAPS else "(" + str(cur) + ")"]
    a[0] += f('Instructions Executed')
    a[1] += f('Thread Instructions Executed')
    a[2] += f('Warp Stall Sampling (All Samples)')
ti = sum(a[0] for a in agg.values())
ts = sum(a[2] for a in agg.values())
print("total warp-instr %.3e  thread-instr %.3e  lanes/instr %.2f  stall samples %d" % (ti, sum(a[1] for a in agg.values()), sum(a[1] for a in agg.values()) / max(ti, 1), ts))
print("%-28s %7s %7s %6s" % ("function", "instr%", "stall%", "lanes"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][2]):
    if a[0] or a[2]:
        print("%-28s %6.1f%% %6.1f%% %6.1f" % (k, 100 * a[0] / ti, 100 * a[2] / max(ts, 1), a[1] / max(a[0], 1)))
