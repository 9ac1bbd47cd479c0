#!/usr/bin/env python
"""Condense an `ncu --csv --metrics ...` launch list into one line per launch: tools/launch_table.py FILE [skip_first_n]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(r[ix['ID']], {'k': r[ix['Kernel Name']], 'b': r[ix['Block Size']], 'g': r[ix['Grid Size']]})
    d[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.defaultdict(float)
for i, d in list(by.items())[skip:]:
    name = d['k'].split('(')[0].replace('void krb::', '').replace('void ', '')
    ms = d.get('gpu__time_duration.sum', 0) / 1e6
    tot[name] += ms
    print('%4s %-34s %-12s %-16s %8.3f ms  R %6.3f W %6.3f GB  fp64 %5.1f  issue %5.1f  warps %5.1f  regs %3d smem %6d' % (
        i, name[:34], d['b'].replace(' ', ''), d['g'].replace(' ', ''), ms, d.get('dram__bytes_read.sum', 0) / 1e9, d.get('dram__bytes_write.sum', 0) / 1e9,
        d.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0), d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0),
        d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0), int(d.get('launch__registers_per_thread', 0)), int(d.get('launch__shared_mem_per_block_dynamic', 0))))
for k, v in tot.items():
    print('total %-40s %9.3f ms' % (k, v))
