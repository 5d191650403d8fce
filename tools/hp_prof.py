#!/usr/bin/env python3
"""Which role bounds the fused inference kernel?  Runs C3 on a library built with -DHP_PROF
(tools/build_exp.sh prof -DHP_PROF; SRCNN_B200_LIB=exp/lib_prof.so python tools/hp_prof.py) and
prints, per warp of CTA (1,1,0), the cycles it spent blocked at its hand-offs against the
cycles of its whole loop.  A role that never waits is the bound.  Kernel-development aid."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
out = subprocess.run([sys.executable, os.path.join(here, "ab_infer.py"), "10"], capture_output=True, text=True)
print(out.stdout.strip().splitlines()[-1])
# the counters are those of the LAST launch of a fresh process: run one more here
sys.path.insert(0, os.path.dirname(here))
sys.argv = [sys.argv[0], "3"]
exec(open(os.path.join(here, "ab_infer.py")).read())
L = C.CDLL(os.environ["SRCNN_B200_LIB"])
buf = (C.c_uint * 128)()
assert L.srcnn_debug_hp_prof(buf) == 0
a = np.array(buf, dtype=np.int64).reshape(32, 4)
roles = [("E1", 0, 8, "bar1 (MMA-1 done)", "(busy) tcgen05.ld"), ("E2", 8, 12, "bar2 (MMA-2 done)", "-"),
         ("E3", 12, 16, "bar3 (MMA-3 done)", "named barrier"), ("IM", 16, 21, "p_free (planes read)", "-"),
         ("I1", 21, 22, "p_full (planes)", "bar2 (D1 free)"), ("I2", 22, 23, "a2_full (E1 done)", "bar3 (D2 free)"),
         ("I3", 23, 24, "a3_full (E2 done)", "d3_free (E3 read D3)")]
for name, w0, w1, n0, n1 in roles:
    for w in range(w0, w1):
        tot, x0, x1, blk = a[w]; nt = 319
        if tot == 0:
            continue
        print("%s warp %2d: %7d cycles, %4d tiles = %6.1f per tile | busy %6.1f | wait %-22s %6.1f | wait %-22s %6.1f" %
              (name, w, tot, nt, tot / nt, (tot - x0 - x1) / nt, n0, x0 / nt, n1, x1 / nt) + (("  | %s %6.1f" % ("MMA issue block" if name[0] == "I" else "tcgen05.wait::st" if name == "E1" else "tcgen05.ld", blk / nt)) if blk else ""))
