#!/usr/bin/env python3
"""Static code size per source function of one kernel, from `nvdisasm -g -c X.cubin` output:
   sass_size.py walk.sass k_walkILb0ELb0    (instructions are 16 bytes; the L1.5 instruction cache holds 32 KB)"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEF = re.compile(r'^(?:static\s+)?(?:template\s*<[^>]*>\s*)?(?:__global__|__device__|MCB_HD|MCB_THD)[^;(]*?\b(\w+)\s*\(')


def function_map(path):
    out, cur, pending = {}, "?", ""
    for i, line in enumerate(open(path), 1):
        text = line.strip()
        if text.startswith("template") and "(" not in text:
            pending = text + " "
            out[i] = cur
            continue
        m = DEF.match(re.sub(r'__launch_bounds__\([^)]*\)', '', pending + text))
        pending = ""
        if m:
            cur = m.group(1)
        out[i] = cur
    return out


maps = {}
want = sys.argv[2]
inside = False
cur = ("?", 0)
agg = collections.Counter()
for line in open(sys.argv[1]):
    if line.startswith("//--------------------- .text."):
        inside = want in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', line):
        f, ln = cur
        base = os.path.basename(f)
        if f not in maps:
            maps[f] = function_map(f) if os.path.exists(f) and f.startswith(ROOT) else None
        name = maps[f].get(ln, base) if maps[f] else "(" + base + ")"
        agg[name] += 1
tot = sum(agg.values())
print("%d instructions = %.1f KB" % (tot, tot * 16 / 1024))
for k, v in agg.most_common(40):
    print("  %-28s %6d  %5.1f KB" % (k, v, v * 16 / 1024))
