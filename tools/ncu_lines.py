#!/usr/bin/env python
"""Warp-stall samples of one kernel by SOURCE LINE: joins `ncu -i X.ncu-rep --page source --csv` (SASS view: one row per instruction,
in address order) with `nvdisasm -g` of the same cubin (line markers).  tools/ncu_lines.py SASS_CSV NVDISASM_TXT MANGLED_KERNEL [top]"""
import collections
import csv
import re
import sys

sass_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst = rows[2:]
# nvdisasm: collect (line, inlined-at chain) per instruction of the kernel's .text section
lines = open(dis).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kern + ':'))
cur = None
per_inst = []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text.'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (m.group(1).split('/')[-1], int(m.group(2)), tuple((a.split('/')[-1], int(b)) for a, b in inl))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        per_inst.append(cur)
assert len(per_inst) >= len(inst), (len(per_inst), len(inst))
agg = collections.Counter()
outer = collections.Counter()
stall = collections.defaultdict(collections.Counter)
total = 0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for k, r in enumerate(inst):
    s = int(r[ix['# Samples']] or 0)
    total += s
    loc = per_inst[k]
    if loc is None:
        continue
    key = (loc[0], loc[1])
    agg[key] += s
    # outermost frame = the line inside the kernel body
    o = loc[2][-1] if loc[2] else key
    outer[o] += s
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            stall[o][c] += v
print('total samples', total)
print('--- by line of the kernel body (outermost frame) ---')
for (f, ln), s in outer.most_common(top):
    st = ', '.join('%s %d' % (c.replace('stall_', ''), v) for c, v in stall[(f, ln)].most_common(3))
    print('%6.2f%%  %s:%d   [%s]' % (100.0 * s / total, f, ln, st))
print('--- by innermost line ---')
for (f, ln), s in agg.most_common(top // 2):
    print('%6.2f%%  %s:%d' % (100.0 * s / total, f, ln))
# --- optional: samples by PHASE = the last kernel-body line (within [lo, hi]) seen in address order; inlined helper code is
#     attributed to the call site that precedes it.  Extra args: lo hi boundary1,boundary2,...
if len(sys.argv) > 7:
    lo, hi = int(sys.argv[5]), int(sys.argv[6])
    bounds = [int(x) for x in sys.argv[7].split(',')]
    cur_line = lo
    ph = collections.Counter()
    phs = collections.defaultdict(collections.Counter)
    for k, r in enumerate(inst):
        loc = per_inst[k]
        if loc and lo <= loc[1] <= hi and loc[0].startswith('front_kernels'):
            cur_line = loc[1]
        b = sum(1 for x in bounds if cur_line >= x)
        s = int(r[ix['# Samples']] or 0)
        ph[b] += s
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                phs[b][c] += v
    print('--- by phase (boundaries %s) ---' % bounds)
    for b in sorted(ph):
        st = ', '.join('%s %.1f%%' % (c.replace('stall_', ''), 100.0 * v / total) for c, v in phs[b].most_common(4))
        print('phase %d: %6.2f%%   [%s]' % (b, 100.0 * ph[b] / total, st))
