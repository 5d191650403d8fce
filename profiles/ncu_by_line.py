#!/usr/bin/env python3
"""Stall samples and executed instructions per CUDA source line of an .ncu-rep captured with
--import-source on (kernel compiled with -lineinfo):  python profiles/ncu_by_line.py <rep> [n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg = collections.OrderedDict()
cur, hdr = None, None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        key = (cur, int(r[0]))
        s, n = int(r[i_s] or 0), int(r[i_i] or 0)
    except ValueError:
        continue
    old = agg.get(key, (r[1].strip()[:84], 0, 0))
    agg[key] = (old[0], old[1] + s, old[2] + n)
tot = sum(v[1] for v in agg.values()) or 1
ti = sum(v[2] for v in agg.values()) or 1
print("total stall samples %d, warp instructions %d" % (tot, ti))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samples %5.1f%% instr  %s:%d  %s" % (100.0 * v[1] / tot, 100.0 * v[2] / ti, k[0], k[1], v[0]))
