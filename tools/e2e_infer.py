#!/usr/bin/env python3
"""e2e timing of srcnn_infer_rows_host on C3 (pinned host buffers, H2D + kernel + D2H)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.load()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import luma_image, make_params
IMG = 4096
rng = np.random.default_rng(1234)
params = make_params(rng, 64, 32, 9, 1, 5)
ctx = pkg.Context(0)
net = pkg.Net(ctx, 64, 32, 9, 1, 5, params)
img = pkg.PinnedBuffer((IMG, IMG)); img.array[:] = luma_image(rng, IMG, IMG)
out = pkg.PinnedBuffer((IMG - 12, IMG - 12))
for _ in range(3): net.infer_rows_host(img.array, IMG, IMG, 0, IMG - 12, out.array)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); net.infer_rows_host(img.array, IMG, IMG, 0, IMG - 12, out.array); ts.append((time.perf_counter() - t0) * 1e3)
print("e2e ms: median %.3f min %.3f  checksum %.9g  (slice=%s, nostream=%s)" % (np.median(ts), min(ts), float(out.array.astype(np.float64).sum()), os.environ.get("SRCNN_STREAM_SLICE"), os.environ.get("SRCNN_NO_STREAMING")))
