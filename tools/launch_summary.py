#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_summary.py profiles/r2m_launches_bench.csv "<command that was profiled>"
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ik, im, iv, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*", "", r[ik].replace("srcnn::", "")).strip()
    name = re.sub(r"^void ", "", name)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("launch list of `%s` under ncu (cold-cache, serialised: compare SHARES)" % sys.argv[2])
print("%-58s %8s %12s %8s %10s" % ("kernel", "launches", "total us", "share", "us/launch"))
for k, v in tot.most_common():
    print("%-58s %8d %12.1f %7.1f%% %10.1f" % (k[:58], cnt[k], v, 100 * v / total, v / cnt[k]))
