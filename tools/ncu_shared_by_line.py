#!/usr/bin/env python3
"""Shared-memory wavefronts per CUDA source line of one kernel of an .ncu-rep (captured with
--import-source on): `L1 Wavefronts Shared` and `L1 Wavefronts Shared Excessive` (= the
wavefronts bank conflicts add) from the source page.
    python tools/ncu_shared_by_line.py <rep> <source file name> [nth capture of that file, default 1] [n lines]
"""
import collections
import csv
import io
import subprocess
import sys

rep, fname = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 1
top = int(sys.argv[4]) if len(sys.argv) > 4 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "File Path" and r[1].endswith(fname)]
start = starts[nth]
end = starts[nth + 1] if nth + 1 < len(starts) else len(rows)
cur, hdr, agg = None, None, collections.OrderedDict()
for r in rows[start:end]:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue

    def g(name):
        try:
            return int(r[hdr.index(name)] or 0)
        except ValueError:
            return 0
    o = agg.setdefault((cur, int(r[0])), [r[1].strip()[:72], 0, 0, 0])
    o[1] += g("L1 Wavefronts Shared")
    o[2] += g("L1 Wavefronts Shared Excessive")
    o[3] += g("Instructions Executed")
print("%s, capture %d of %s: shared wavefronts %d, of which excessive (bank conflicts) %d; warp instructions %d" % (
    rep.split("/")[-1], nth, fname, sum(v[1] for v in agg.values()), sum(v[2] for v in agg.values()),
    sum(v[3] for v in agg.values())))
print("%10s %10s %11s  line" % ("wavefronts", "excessive", "warp instr"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%10d %10d %11d  %s:%d  %s" % (v[1], v[2], v[3], k[0], k[1], v[0]))
