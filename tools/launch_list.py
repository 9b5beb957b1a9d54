#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and total time per kernel.
usage: launch_list.py launches.csv > out.txt"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    a = agg.setdefault(r[ix["Kernel Name"]][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in agg.items():
    print("%-70s   n=%3d total %10.1f us (%4.1f%%)" % (k, n, t, 100 * t / tot))
