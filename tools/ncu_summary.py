#!/usr/bin/env python3
"""Summarise ncu outputs: `ncu_summary.py launches <csv>` (gpu__time_duration launch list) or
`ncu_summary.py raw <csv from ncu -i X.ncu-rep --page raw --csv>`."""
import collections
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__icc_request_hit_rate.pct',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed', 'sm__icc_requests.sum.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split('(')[0].replace('<unnamed>::', '')
        if 'cub::' in name:
            name = 'cub::' + name.split('cub::')[1].split('<')[0]
        v = float(r[vi].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(v[1] for v in agg.values())
    print("%-28s %6s %12s %10s %8s" % ("kernel", "n", "total_ms", "max_us", "share"))
    for k, (n, v, m) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-28s %6d %12.3f %10.1f %7.1f%%" % (k, n, v / 1e6, m / 1e3, 100 * v / tot))
    print("%-28s %6d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot / 1e6))


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('---- ' + r[ki].split('(')[0].replace('<unnamed>::', ''))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-78s %18s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
