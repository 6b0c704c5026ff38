#!/usr/bin/env python
"""Key per-launch metrics of an .ncu-rep as a markdown table:
   python profiles/ncu_metrics.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
want = [('Kernel Name', 'kernel'), ('Grid Size', 'grid'), ('gpu__time_duration.sum', 'us'),
        ('dram__bytes_read.sum', 'dram rd MB'), ('dram__bytes_write.sum', 'dram wr MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor %'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
        ('launch__registers_per_thread', 'regs'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ %')]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print('| ' + ' | '.join(n for _, n in idx) + ' |')
print('|' + '---|' * len(idx))
for r in data:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == 'kernel':
            v = v.split('(')[0].replace('void ', '').replace('vdm::<unnamed>::', '')[:60]
        cells.append(v)
    print('| ' + ' | '.join(cells) + ' |')
