"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares.
Usage: python tools/summarize_launches.py launches.csv [skip_first_n] [count]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v / 1000.0 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1000.0
    name = r['Kernel Name']
    short = name.split('(')[0].split('<')[0].split('::')[-1].replace('void ', '').strip()
    rows.append((int(r['ID']), short, us, r['Grid Size'], r['Block Size']))
rows = rows[skip:skip + count]
tot = sum(r[2] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for _, n, us, _, _ in rows:
    agg[n][0] += 1
    agg[n][1] += us
print(f'{len(rows)} launches, total {tot / 1000:.3f} ms (cold-cache, serialised: compare shares)')
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n:42s} n={c:4d}  {us / 1000:8.3f} ms  {100 * us / tot:5.1f}%  avg {us / c:8.1f} us')
if '--list' in sys.argv:
    for i, n, us, g, b in rows:
        print(f'{i:5d} {n:36s} {us:9.1f} us grid={g} block={b}')
