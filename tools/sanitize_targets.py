#!/usr/bin/env python3
"""One small pass over every tensor-core kernel family, sized for `compute-sanitizer --tool
memcheck` (run plainly first; see tools/gpu_run5.sh): training chunks of 9-1-5 and 9-5-5 at
the smallest batch that takes the tensor-core paths, with ragged widths, the fused inference
kernels on an image with ragged strips / bands, and the two host-buffer pipelines."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
from helpers import luma_image, make_params, patches  # noqa: E402

pkg = _pkg.load()
rng = np.random.default_rng(11)
ctx = pkg.Context(0, profile=True)
for cfg, S in (((64, 32, 9, 1, 5), 150), ((64, 32, 9, 5, 5), 197), ((64, 32, 9, 5, 5), 333)):
    net = pkg.Net(ctx, *cfg, make_params(rng, *cfg))
    x, gt = patches(rng, S, 33, 33)
    mi, mg = ctx.upload(x), ctx.upload(gt)
    work = ctx.alloc(net.train_workspace_bytes(33, 33, S))
    net.train_chunk(mi, mg, 33, 33, S, work)
    net.update_all(S, 0.9, 0.001, [1e-4, 1e-4, 1e-5])
    ctx.block()
    assert all(np.isfinite(g).all() for g in net.grads().values())
    for m in (mi, mg, work):
        ctx.release(m)
    print("train", cfg, S, "ok")
for cfg, (w, h, s) in (((64, 32, 9, 1, 5), (301, 277, 1)), ((128, 64, 9, 1, 5), (260, 141, 3))):
    net = pkg.Net(ctx, *cfg, make_params(rng, *cfg))
    x = np.stack([luma_image(rng, h, w) for _ in range(s)])
    mi, mo = ctx.upload(x), ctx.alloc(4 * s * (w - 12) * (h - 12))
    net.forward_fused(mi, mo, w, h, s)
    y = ctx.read(mo, (s, h - 12, w - 12))
    assert np.isfinite(y).all()
    hin, hout = pkg.PinnedBuffer(x.shape), pkg.PinnedBuffer(y.shape)
    hin.array[:] = x
    if s == 1:
        net.infer_rows_host(hin.array, w, h, 0, h - 12, hout.array)
    else:
        net.infer_frames_host(hin.array, w, h, hout.array)
    assert np.array_equal(hout.array, y)
    hin.free()
    hout.free()
    print("infer", cfg, (w, h, s), "ok")
print("kernels:", sorted(k for k, (ns, n) in ctx.profile().items() if n))
ctx.close()
print("ok")
