#!/usr/bin/env python
"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name for ONE call of the profiled script:
usage: python tools/launch_agg.py FILE.csv MARKER [occurrence]   (the call spans MARKER's occurrence-th .. next launch)"""
import collections, csv, re, sys
f, marker = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rows = []
lines = [l for l in open(f) if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        v = float(x["Metric Value"].replace(",", "")); u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        rows.append((x["Kernel Name"], v, x.get("Grid Size")))
starts = [i for i, r in enumerate(rows) if marker in r[0]]
s, e = starts[occ], starts[occ + 1]
agg = collections.OrderedDict()
for r in rows[s:e]:
    k = re.sub(r"\(.*", "", r[0])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += r[1]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:9.1f} us {a[0]:4d}  {k}")
print(f"{sum(a[1] for a in agg.values()):9.1f} us {e - s:4d}  total")
