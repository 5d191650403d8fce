"""Shared input generators for the parity tests (seeded, synthetic)."""
import numpy as np


def make_params(rng, n1, n2, f1, f2, f3, sd=None, bias_sd=0.01):
    """Random parameters.  With sd=None the scale is chosen so that activations stay O(1)
    through the three layers (He-style), which keeps ReLUs alive and makes an absolute 1e-4
    tolerance on the output meaningful; the reference draws N(0, 0.001..0.005)
    (example_config.json:12-29) which gives near-zero outputs."""
    k = [1, n1, n2]
    n = [n1, n2, 1]
    f = [f1, f2, f3]
    p = {}
    for l in range(3):
        s = sd if sd is not None else np.sqrt(2.0 / (f[l] * f[l] * k[l]))
        p["w%d" % (l + 1)] = rng.normal(0, s, f[l] * f[l] * k[l] * n[l]).astype(np.float32)
        p["b%d" % (l + 1)] = rng.normal(0, bias_sd, n[l]).astype(np.float32)
    p["b3"] = (p["b3"] + 0.2).astype(np.float32)
    return p


def luma_image(rng, h, w):
    """Smooth-ish luma in [0,1): low-pass filtered uniform noise."""
    x = rng.uniform(0, 1, (h + 4, w + 4)).astype(np.float32)
    y = (x[:-4, :-4] + x[2:-2, 2:-2] + x[4:, 4:] + x[:-4, 4:] + x[4:, :-4]) / 5.0
    return np.ascontiguousarray(y, np.float32)


def patches(rng, n, w, h):
    """Ground truth U[0,1); input = truth + N(0,0.05^2) clipped, minus its mean."""
    gt = rng.uniform(0, 1, (n, h, w)).astype(np.float32)
    x = np.clip(gt + rng.normal(0, 0.05, gt.shape), 0, 1).astype(np.float32)
    x -= x.mean(axis=(1, 2), keepdims=True)
    return np.ascontiguousarray(x, np.float32), gt


def psnr(sse, count):
    """PSNR = 10 log10(1/MSE) from the squared-error sum (not in the reference)."""
    return 10.0 * np.log10(count / max(sse, 1e-300))
