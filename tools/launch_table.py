"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
        rows.append((r['Kernel Name'], us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0.0, 0])
for k, u in rows:
    agg[k][0] += u
    agg[k][1] += 1
print('%d launches, %.1f us serialised' % (len(rows), tot))
for k, (u, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%6.2f%% %10.1f us %4d  %s' % (100 * u / tot, u, n, k[:110]))
