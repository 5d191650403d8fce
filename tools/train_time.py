#!/usr/bin/env python3
"""A/B timing of one training chunk (C2 shape: 2048 patches 33x33, 9-1-5 64/32; C4 with "955").
    SRCNN_B200_LIB=exp/lib_x.so python tools/train_time.py [S [955]]
Prints ms per chunk (CUDA events on the context stream, median of 10) and a checksum of the
gradients.  Not part of the bench contract."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.load()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_params, patches
rng = np.random.default_rng(7)
F2 = 5 if (len(sys.argv) > 2 and sys.argv[2] == "955") else 1   # "955": the 9-5-5 net of C4
params = make_params(rng, 64, 32, 9, F2, 5)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = pkg.Context(0, stream=stream.cuda_stream)
    net = pkg.Net(ctx, 64, 32, 9, F2, 5, params)
    x, gt = patches(rng, S, 33, 33)
    mi, mg = ctx.upload(x), ctx.upload(gt)
    work = ctx.alloc(net.train_workspace_bytes(33, 33, S))
    for _ in range(3):
        net.train_chunk(mi, mg, 33, 33, S, work)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        net.train_chunk(mi, mg, 33, 33, S, work)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    g = net.grads()
    print("%s: median %.4f ms  min %.4f ms per chunk of %d; grad checksum %.6e" % (
        os.path.basename(pkg.LIB_PATH), float(np.median(ts)), min(ts), S,
        sum(float(np.abs(v).sum()) for v in g.values())))
