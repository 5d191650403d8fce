#!/usr/bin/env python3
"""One-off timing of the other named inference configs (not bench lines): C5's network
(9-1-5 n1=128 n2=64) on 1920x1080 frames, and the 64/32 network on the same frames."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.load()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import luma_image, make_params
rng = np.random.default_rng(5)
W, H = 1920, 1080
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = pkg.Context(0, stream=stream.cuda_stream)
    for (n1, n2, S) in ((128, 64, 4), (64, 32, 4), (64, 32, 32)):
        net = pkg.Net(ctx, n1, n2, 9, 1, 5, make_params(rng, n1, n2, 9, 1, 5))
        x = np.stack([luma_image(rng, H, W) for _ in range(min(S, 4))] * (S // min(S, 4)))
        mi, mo = ctx.upload(x), ctx.alloc(4 * S * (W - 12) * (H - 12))
        for _ in range(2):
            net.forward_fused(mi, mo, W, H, S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            net.forward_fused(mi, mo, W, H, S)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("9-1-5 n1=%d n2=%d, %d frames %dx%d: %.3f ms per call, %.3f ms per frame, %.0f MPix/s" %
              (n1, n2, S, W, H, ms, ms / S, S * W * H / 1e6 / (ms / 1e3)))
        ctx.release(mi); ctx.release(mo)
