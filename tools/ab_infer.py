#!/usr/bin/env python3
"""A/B timing of the fused inference launch (C3: 4096x4096, 9-1-5 64/32) for kernel work.
    SRCNN_B200_LIB=exp/lib_x.so python tools/ab_infer.py [steps] [size]
Prints per-launch ms (median / min over `steps` launches, CUDA events on the launching stream,
4 rotating buffer pairs) and a checksum of the output, so variants can be compared for speed
and for equality of results.  Not part of the bench contract."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg  # noqa: E402

pkg = _pkg.load()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
IMG = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import luma_image, make_params  # noqa: E402

rng = np.random.default_rng(1234)
params = make_params(rng, 64, 32, 9, 1, 5)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = pkg.Context(0, stream=stream.cuda_stream)
    net = pkg.Net(ctx, 64, 32, 9, 1, 5, params)
    img = luma_image(rng, IMG, IMG)
    R = 4
    ins = [ctx.upload(np.roll(img, i, axis=1)) for i in range(R)]
    outs = [ctx.alloc(4 * (IMG - 12) * (IMG - 12)) for _ in range(R)]
    for i in range(3):
        net.forward_fused(ins[i % R], outs[i % R], IMG, IMG, 1)
    torch.cuda.synchronize()
    ts = []
    for i in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        net.forward_fused(ins[i % R], outs[i % R], IMG, IMG, 1)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = ctx.read(outs[0], (IMG - 12, IMG - 12))
    ts = np.array(ts)
    print("%s: median %.4f ms  min %.4f ms  max %.4f  checksum %.9g  absmax %.6g" %
          (os.path.basename(os.environ.get("SRCNN_B200_LIB", "default")), np.median(ts), ts.min(),
           ts.max(), float(out.astype(np.float64).sum()), float(np.abs(out).max())))
