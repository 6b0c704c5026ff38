#!/usr/bin/env python
"""profiles/ncu_top_launch.json from an `ncu --set full` report of profiles/gemm_ncu_probe.py: DRAM bytes, duration and
tensor-pipe utilisation of the heaviest launch shape of the C2 step (conv3x3 64x64 256 -> 128, transposed-role halo
kernel).  bench.py reads the file for `roofline.traffic`.
   python profiles/make_top_launch_json.py gpurun_out/ncu_gemm_r2.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
col = {k: hdr.index(k) for k in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                                 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed') if k in hdr}
units = rows[1]


def val(r, k):
    v = float(r[col[k]].replace(',', ''))
    u = units[col[k]]
    return v * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'us': 1.0, 'ms': 1e3, 'ns': 1e-3, '%': 1.0}.get(u, 1.0)


top = [r for r in data if 'gemm_tc_halo_t_kernel' in r[col['Kernel Name']]][0]      # first probe shape: conv64 256 -> 128
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
build = subprocess.run(['git', '-C', root, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
js = {'kernel': top[col['Kernel Name']].split('(')[0][-80:], 'shape': 'conv3x3 64x64 256->128, M=655360 N=128 K=2304, bf16 out + statistics',
      'us_under_ncu': val(top, 'gpu__time_duration.sum'), 'dram_bytes_read': val(top, 'dram__bytes_read.sum'),
      'dram_bytes_write': val(top, 'dram__bytes_write.sum'),
      'tensor_pipe_pct': val(top, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
      'build': build, 'source': f'ncu --set full --clock-control none python profiles/gemm_ncu_probe.py -> {os.path.basename(rep)}'}
with open(os.path.join(root, 'profiles', 'ncu_top_launch.json'), 'w') as f:
    json.dump(js, f, indent=1)
print(json.dumps(js, indent=1))
