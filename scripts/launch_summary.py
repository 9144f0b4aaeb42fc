#!/usr/bin/env python3
"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_summary.py launches.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
    a[0] += 1; a[1] += float(r[ix["Metric Value"]])
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_ns,share,avg_ns,grid,block")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{a[0]},{a[1]:.0f},{a[1]/tot:.4f},{a[1]/a[0]:.0f},"{a[2]}","{a[3]}"')
