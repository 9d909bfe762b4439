"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/agg_launches.py launches.csv [first_marker_kernel]  (the last segment between two
launches of the marker kernel, default adam_kernel, is summarised)"""
import collections
import csv
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else 'adam_kernel'
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if marker in x['Kernel Name']]
seg = rows[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for x in seg:
    n = x['Kernel Name'].split('(')[0]
    v = float(x['Metric Value'].replace(',', ''))
    u = x['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(seg)} launches")
print("| share | launches | total us | avg us | kernel |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"| {100 * t / tot:.1f} % | {c} | {t:.1f} | {t / c:.1f} | `{n.replace('void ', '')[:90]}` |")
