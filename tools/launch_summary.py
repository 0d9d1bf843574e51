"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        h, start = r, i + 1
        break
ik, iv, iu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
t, n = collections.Counter(), collections.Counter()
for r in rows[start:]:
    if len(r) <= iv:
        continue
    name = r[ik].split('(')[0].split('::')[-1]
    v = float(r[iv].replace(',', ''))
    ms = v / 1e6 if r[iu] == 'ns' else v / 1e3 if r[iu] == 'us' else v
    t[name] += ms
    n[name] += 1
tot = sum(t.values())
for k, v in t.most_common():
    print(f"{k:45s} n={n[k]:5d} total={v:9.3f} ms  avg={v / n[k] * 1e3:9.1f} us  share={v / tot * 100:5.1f}%")
print(f"sum {tot:.3f} ms")
