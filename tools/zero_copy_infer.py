#!/usr/bin/env python3
"""Experiment: C3 end to end with the fused kernel reading / writing PINNED HOST memory directly
(unified addressing: a cudaHostAlloc pointer is a device pointer) instead of staged copies.
    python tools/zero_copy_infer.py
Prints wall-clock ms per image (srcnn_block after each) for
  staged    srcnn_infer_rows_host (the product path: pipelined sub-band copies)
  zc-both   one launch, input and output in host memory
  zc-out    whole-image H2D copy, one launch writing host memory
Kernel-development aid, not part of the bench contract."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

pkg = _pkg.load()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import luma_image, make_params  # noqa: E402

IMG = 4096
O = IMG - 12
rng = np.random.default_rng(1234)
params = make_params(rng, 64, 32, 9, 1, 5)
ctx = pkg.Context(0)
net = pkg.Net(ctx, 64, 32, 9, 1, 5, params)
R = 3
imgs = [pkg.PinnedBuffer((IMG, IMG)) for _ in range(R)]
outs = [pkg.PinnedBuffer((O, O)) for _ in range(R)]
base = luma_image(rng, IMG, IMG)
for i, b in enumerate(imgs):
    b.array[:] = np.roll(base, i, axis=1)
h_in = [ctx.wrap(b.ptr, b.nbytes) for b in imgs]
h_out = [ctx.wrap(b.ptr, b.nbytes) for b in outs]
d_in = ctx.alloc(4 * IMG * IMG)


def staged(i):
    net.infer_rows_host(imgs[i].array, IMG, IMG, 0, O, outs[i].array)


def zc_both(i):
    net.forward_fused(h_in[i], h_out[i], IMG, IMG, 1)
    ctx.block()


def zc_out(i):
    ctx.write(d_in, imgs[i].array, block=False)
    net.forward_fused(d_in, h_out[i], IMG, IMG, 1)
    ctx.block()


for name, fn in (("staged", staged), ("zc-both", zc_both), ("zc-out", zc_out)):
    for i in range(3):
        fn(i % R)
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        fn(i % R)
        ts.append((time.perf_counter() - t0) * 1e3)
    cs = float(outs[(12 - 1) % R].array.astype(np.float64).sum())
    print("%-8s median %.3f ms  min %.3f ms  checksum %.9g" % (name, np.median(ts), min(ts), cs))
