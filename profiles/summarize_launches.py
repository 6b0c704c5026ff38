#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
   python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    m = re.search(r'(\w+_kernel|\w+)\s*(<[^(]*>)?\s*\(', r[ki])
    name = m.group(1) + (m.group(2) or '') if m else r[ki][:60]
    name = re.sub(r'\(anonymous namespace\)::|vdm::|void ', '', name)
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f'| kernel | launches | total us | share |\n|---|---:|---:|---:|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |')
print(f'| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |')
