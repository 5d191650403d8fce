#!/usr/bin/env python3
"""A short run of the kernels the bench is made of, for `ncu` (launch list and --set full
captures): C3 fused inference, one C2 training chunk of 2048 patches, one group of 4 C5 frames
and (with `c4` in argv) one C4 training chunk.  Every kernel runs twice: the first launch is the
warm-up, the second the one to read."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
from helpers import luma_image, make_params, patches  # noqa: E402

pkg = _pkg.load()
rng = np.random.default_rng(7)
what = set(sys.argv[1:]) or {"c3", "c2", "c5"}
ctx = pkg.Context(0)
if "c3" in what:
    net = pkg.Net(ctx, 64, 32, 9, 1, 5, make_params(rng, 64, 32, 9, 1, 5))
    x = luma_image(rng, 4096, 4096)
    mi, mo = ctx.upload(x), ctx.alloc(4 * 4084 * 4084)
    for _ in range(2):
        net.forward_fused(mi, mo, 4096, 4096, 1)
    ctx.block()
    ctx.release(mi)
    ctx.release(mo)
for key, cfg in (("c2", (64, 32, 9, 1, 5)), ("c4", (64, 32, 9, 5, 5))):
    if key in what:
        net = pkg.Net(ctx, *cfg, make_params(rng, *cfg))
        S = 2048
        x, gt = patches(rng, S, 33, 33)
        mi, mg = ctx.upload(x), ctx.upload(gt)
        work = ctx.alloc(net.train_workspace_bytes(33, 33, S))
        for _ in range(2):
            net.train_chunk(mi, mg, 33, 33, S, work)
        net.update_all(S, 0.9, 0.001, [1e-4, 1e-4, 1e-5])
        ctx.block()
        for m in (mi, mg, work):
            ctx.release(m)
if "c5" in what:
    net = pkg.Net(ctx, 128, 64, 9, 1, 5, make_params(rng, 128, 64, 9, 1, 5))
    x = np.stack([luma_image(rng, 1080, 1920) for _ in range(4)])
    mi, mo = ctx.upload(x), ctx.alloc(4 * 4 * 1908 * 1068)
    for _ in range(2):
        net.forward_fused(mi, mo, 1920, 1080, 4)
    ctx.block()
ctx.close()
print("ok")
