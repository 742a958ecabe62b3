#!/usr/bin/env python3
"""Instruction and stall-sample shares per code region of k_step from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys


def region(f, l):
    l = int(l)
    if f == 'mcb_tables.h':
        return 'lookup: union-grid search'
    if f == 'mcb_device.cuh':
        for hi, name in ((170, 'lookup: micro_xs rows'), (200, 'lookup: macro sums'), (245, 'select nuclide'), (275, 'watt'),
                         (295, 'isotropic / dist1'), (340, 'scatter_sample'), (10**9, 'surface_intersect')):
            if l <= hi:
                return name
    if f == 'mcb_physics.h':
        for hi, name in ((62, 'rng'), (100, 'algorithm (quad, interp)'), (195, 'surface eval/distance'), (213, 'search_cell'),
                         (240, 'scatter_direction'), (10**9, 'physics other')):
            if l <= hi:
                return name
    if f == 'mcb_kernels.cu':
        for hi, name in ((71, 'block_reserve'), (150, 'tally score'), (196, 'ev_flight'), (241, 'ev_collide_pre'),
                         (282, 'ev_collide_bank'), (303, 'ev_collide_scatter (self)'), (370, 'ev_cross'), (10**9, 'k_step body')):
            if l <= hi:
                return name
    return f


rows = list(csv.reader(open(sys.argv[1])))
cur = None
hdr = None
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
    a = agg[region(cur, r[0])]
    a[0] += f('Instructions Executed')
    a[1] += f('Thread Instructions Executed')
    a[2] += f('Warp Stall Sampling (All Samples)')
ti = sum(a[0] for a in agg.values())
ts = sum(a[2] for a in agg.values())
print("%-28s %7s %7s %6s" % ("region", "instr%", "stall%", "lanes"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][2]):
    print("%-28s %6.1f%% %6.1f%% %6.1f" % (k, 100 * a[0] / ti, 100 * a[2] / ts, a[1] / max(a[0], 1)))
