"""Averages the per-launch DRAM traffic of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum --csv` capture over the LAST forward in the file and writes profiles/conv_gemm_traffic.json
(read by bench.py for roofline.traffic).  Usage: python tools/summarize_traffic.py capture.csv"""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
per = defaultdict(dict)
order = []
for r in csv.DictReader(lines):
    i = int(r['ID'])
    if i not in per:
        order.append(i)
    name = r['Kernel Name'].split('(')[0].split('<')[0].split('::')[-1].replace('void ', '').strip()
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if r['Metric Name'].startswith('dram__bytes'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    else:
        v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3}[u]
    per[i]['name'] = name
    per[i][r['Metric Name']] = v
rows = [per[i] for i in order]
half = rows[len(rows) // 2:]      # second (steady-state) forward
out = {}
for key, label in (('conv_gemm', 'conv_gemm'), ('groupnorm_apply', 'groupnorm_apply')):
    sel = [r for r in half if r['name'].startswith(key)]
    if not sel:
        continue
    rd = sum(r['dram__bytes_read.sum'] for r in sel)
    wr = sum(r['dram__bytes_write.sum'] for r in sel)
    us = sum(r['gpu__time_duration.sum'] for r in sel)
    out[label] = {'launches': len(sel), 'dram_read_bytes': rd, 'dram_write_bytes': wr,
                  'dram_bytes_per_launch': (rd + wr) / len(sel), 'us_total_cold': us,
                  'dram_gbs_cold': (rd + wr) / us / 1e3}
res = {'source': os.path.basename(path), 'workload': 'CIFAR-10 UNet forward, batch 256 (ncu, cold cache per launch)',
       'dram_bytes_per_launch': out.get('conv_gemm', {}).get('dram_bytes_per_launch'), 'kernels': out}
with open(os.path.join(ROOT, 'profiles', 'conv_gemm_traffic.json'), 'w') as f:
    json.dump(res, f, indent=1)
print(json.dumps(res))
