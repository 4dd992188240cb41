#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]].split("(")[0][:90]
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{t/1e3:10.3f} ms {100*t/tot:5.1f}%  x{n:4d}  {k}")
