#!/usr/bin/env python
"""Top warp-stall source lines / SASS of an ncu report with source (`--import-source on`).
usage: tools/ncu_src_stalls.py REPORT.ncu-rep [N]  -> per CUDA source line: samples and dominant stall reasons"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.OrderedDict(); total = 0
cur_file = ""
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ci["# Samples"]] or 0)
    except ValueError: continue
    key = (r[0], r[1].strip()[:110])
    a = agg.setdefault(key, [0, collections.Counter(), 0])
    a[0] += n; total += n
    try: a[2] += int(r[ci["Instructions Executed"]] or 0)
    except ValueError: pass
    for s in stall_cols:
        try: a[1][s] += int(r[ci[s]] or 0)
        except ValueError: pass
print("total samples", total)
for (ln, src), (n, st, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3) if v)
    print(f"{n:6d} {100.0 * n / max(total, 1):5.1f}%  L{ln:>5s} inst={ie:8d} [{top}]  {src}")
