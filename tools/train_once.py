#!/usr/bin/env python3
"""Runs a few training chunks of C2 (for profiling the training kernels under ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.load()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_params, patches
rng = np.random.default_rng(7)
params = make_params(rng, 64, 32, 9, 1, 5)
ctx = pkg.Context(0)
net = pkg.Net(ctx, 64, 32, 9, 1, 5, params)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x, gt = patches(rng, S, 33, 33)
mi, mg = ctx.upload(x), ctx.upload(gt)
work = ctx.alloc(net.train_workspace_bytes(33, 33, S))
for _ in range(3):
    net.train_chunk(mi, mg, 33, 33, S, work)
ctx.block()
print("ok")
