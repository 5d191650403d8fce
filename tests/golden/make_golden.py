#!/usr/bin/env python3
"""Generates the committed fixtures under tests/golden/.  Run in the dev container
(needs /root/reference and oracle/_ref/libsrcnn_ref.so):

    python tests/golden/make_golden.py

Two kinds of fixture:

1. *Reference test data*, transcribed mechanically from the reference's own specs:
     layer_cases.json        <- test/data/test_cases.json                     (forward, 3 shapes)
     layer_deltas_case.json  <- test/specs/LayerDeltasTest.cpp:33-126         (deltas)
     backprop_case.json      <- test/specs/BackpropagationTest.cpp:31-90      (gW with 1.5 pre-fill, gB)
2. *Reference outputs*: the reference's OWN kernels (compiled by g++ through
   oracle/cl_shim.hpp -> oracle/_ref/libsrcnn_ref.so) run on seeded random inputs:
     ref_kernels_small.npz   per-kernel, multi-sample (S>1) cases the reference's tests never pin
     ref_train_chain.npz     forward -> backward -> update, 2 epochs x 2 chunks, tiny 3-1-3 net,
                             weight decay on (the reference leaves the decay term untested)
     ref_train_915.npz       same chain on a 9-1-5 n1=8 n2=4 net with 33x33 patches (outputs only:
                             parameters after 2 epochs + final SSE)

     ref_train_c2_epoch.npz  BASELINE config C2 exactly as bench.py trains it: 9-1-5 n1=64 n2=32,
                             4096 patches 33x33, chunks of 2048, momentum 0.9, decay 1e-3,
                             lr 1e-4/1e-4/1e-5; parameters after epochs 1 and 2 (inputs are
                             regenerated from the seed by tests/helpers.py)
     ref_train_c4_epoch.npz  the 9-5-5 network of config C4 on 512 patches, chunks of 256
3. *Reference image fixtures* for the luma kernels, decoded with the reference's own vendored
   stb_image (libs/include/stb, compiled here into a throw-away decoder: PIL's JPEG decoder
   differs from it by up to 2 levels, which would break SwapLumaTest's exact comparison):
     color_grid_5x5.rgba                 <- test/data/color_grid.png        (ExtractLumaTest.cpp)
     color_grid2_32x32.rgba              <- test/data/color_grid2.jpg       (SwapLumaTest.cpp)
     color_grid2_luma_swapped_32x32.rgb  <- test/data/color_grid2_luma_swapped.png
     luma_goldens.json                   <- test/specs/ExtractLumaTest.cpp:24-28 (25 values)

Nothing under /root/reference is needed at test time.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.loader import NetState, Oracle  # noqa: E402


def c_array(src, name):
    """Pull `name[...] = { ... }` / `name = { ... }` float initialiser out of C++ source."""
    m = re.search(r"\b" + re.escape(name) + r"\s*(\[[^\]]*\])?\s*=\s*\{(.*?)\};", src, re.S)
    assert m, name
    body = re.sub(r"//[^\n]*", "", m.group(2))
    vals = re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", body)
    return [float(v) for v in vals]


def transcribe_reference_specs():
    with open(os.path.join(REF, "test/data/test_cases.json")) as fh:
        cases = json.load(fh)
    with open(os.path.join(HERE, "layer_cases.json"), "w") as fh:
        json.dump({"source": "test/data/test_cases.json", "cases": cases}, fh, indent=1)

    src = open(os.path.join(REF, "test/specs/LayerDeltasTest.cpp")).read()
    d = {
        "source": "test/specs/LayerDeltasTest.cpp:33-126,141-193",
        "note": "prev layer n=2 on 5x5; next layer f=3 n=3 on 3x3; "
                "layer_output = max(input_x, 0) (test/TestCase.cpp:15)",
        "n_curr": 2, "f_next": 3, "n_next": 3, "out_w": 5, "out_h": 5,
        "input_x": c_array(src, "input_x"),
        "weights": c_array(src, "weights"),
        "deltas": c_array(src, "deltas"),
        "expected_output": c_array(src, "expected_output"),
    }
    assert len(d["input_x"]) == 50 and len(d["weights"]) == 54
    assert len(d["deltas"]) == 27 and len(d["expected_output"]) == 50
    with open(os.path.join(HERE, "layer_deltas_case.json"), "w") as fh:
        json.dump(d, fh, indent=1)

    src = open(os.path.join(REF, "test/specs/BackpropagationTest.cpp")).read()
    b = {
        "source": "test/specs/BackpropagationTest.cpp:31-90,135-171",
        "note": "k=2 n=3 f=3, 5x5 input, 3x3 deltas, grad_w pre-filled with 1.5 (accumulation)",
        "k": 2, "n": 3, "f": 3, "out_w": 3, "out_h": 3, "grad_w_init": 1.5,
        "input": c_array(src, "input"),
        "deltas": c_array(src, "deltas"),
        "expected_weights": c_array(src, "expected_weights"),
        "expected_bias": c_array(src, "expected_bias"),
    }
    assert len(b["input"]) == 50 and len(b["deltas"]) == 27
    assert len(b["expected_weights"]) == 54 and len(b["expected_bias"]) == 3
    with open(os.path.join(HERE, "backprop_case.json"), "w") as fh:
        json.dump(b, fh, indent=1)


def copy_config_fixtures():
    """test/data/config*.json -> tests/golden/ref_config/ (reference: test/specs/ConfigTest.cpp)"""
    import shutil
    dst = os.path.join(HERE, "ref_config")
    os.makedirs(dst, exist_ok=True)
    for f in ("config.json", "config_invalid_val.json", "config_non_parseable.json"):
        shutil.copy(os.path.join(REF, "test/data", f), os.path.join(dst, f))
    shutil.copy(os.path.join(REF, "example_config.json"), os.path.join(dst, "example_config.json"))


def make_params(rng, n1, n2, f1, f2, f3, sd=0.05, bias_sd=0.01):
    return {
        "w1": rng.normal(0, sd, f1 * f1 * 1 * n1).astype(np.float32),
        "b1": rng.normal(0, bias_sd, n1).astype(np.float32),
        "w2": rng.normal(0, sd, f2 * f2 * n1 * n2).astype(np.float32),
        "b2": rng.normal(0, bias_sd, n2).astype(np.float32),
        "w3": rng.normal(0, sd, f3 * f3 * n2).astype(np.float32),
        "b3": rng.normal(0.05, bias_sd, 1).astype(np.float32),
    }


def reference_kernel_outputs(ref):
    rng = np.random.default_rng(20261018)
    out = {}
    # forward, S=3, k=3 n=5 f=3, 7x6 (generic shape: exercises the run-time instantiation)
    S, k, n, f, w, h = 3, 3, 5, 3, 7, 6
    x = rng.normal(0, 1, (S, h, w, k)).astype(np.float32)
    W = rng.normal(0, 0.3, f * f * k * n).astype(np.float32)
    B = rng.normal(0, 0.1, n).astype(np.float32)
    out.update(fw_x=x, fw_W=W, fw_B=B, fw_shape=np.array([S, k, n, f, w, h]),
               fw_relu=ref.forward(x, W, B, k, n, f, False, w, h, S),
               fw_lin=ref.forward(x, W, B, k, n, f, True, w, h, S))
    # last layer delta + squared error, S=2, 5x4 result in 9x8 ground truth
    S, aw, ah, pad = 2, 5, 4, 4
    gt = rng.uniform(0, 1, (S, ah + pad, aw + pad)).astype(np.float32)
    algo = rng.normal(0.3, 0.5, (S, ah, aw)).astype(np.float32)
    out.update(ll_gt=gt, ll_algo=algo,
               ll_delta=ref.last_layer_delta(gt, algo, aw + pad, ah + pad, aw, ah, S),
               ll_sse=np.float64(ref.squared_error(gt, algo, aw + pad, ah + pad, aw, ah, S)))
    # deltas, S=2, n_curr=4, next f=3 n=5, out 6x5
    S, nc, fn, nn, ow, oh = 2, 4, 3, 5, 6, 5
    dn = rng.normal(0, 1, (S, oh - fn + 1, ow - fn + 1, nn)).astype(np.float32)
    lo = rng.normal(0, 1, (S, oh, ow, nc)).astype(np.float32)
    lo = np.maximum(lo, 0).astype(np.float32)
    W = rng.normal(0, 0.3, fn * fn * nc * nn).astype(np.float32)
    out.update(dl_next=dn, dl_out=lo, dl_W=W, dl_shape=np.array([S, nc, fn, nn, ow, oh]),
               dl_result=ref.deltas(dn, lo, W, nc, fn, nn, ow, oh, S))
    # backpropagate, S=3, k=3 n=4 f=3, out 5x4, accumulators pre-filled
    S, k, n, f, ow, oh = 3, 3, 4, 3, 5, 4
    d = rng.normal(0, 1, (S, oh, ow, n)).astype(np.float32)
    li = rng.normal(0, 1, (S, oh + f - 1, ow + f - 1, k)).astype(np.float32)
    gw = rng.normal(0, 1, f * f * k * n).astype(np.float32)
    gb = rng.normal(0, 1, n).astype(np.float32)
    out.update(bp_d=d, bp_in=li, bp_gw0=gw.copy(), bp_gb0=gb.copy(),
               bp_shape=np.array([S, k, n, f, ow, oh]))
    ref.backpropagate(d, li, gw, gb, n, k, f, ow, oh, S)
    out.update(bp_gw=gw, bp_gb=gb)
    # update_params with non-zero weight decay (untested by the reference)
    ws, bs = 37, 5
    wv = rng.normal(0, 1, ws).astype(np.float32)
    bv = rng.normal(0, 1, bs).astype(np.float32)
    gwv = rng.normal(0, 3, ws).astype(np.float32)
    gbv = rng.normal(0, 3, bs).astype(np.float32)
    pw = rng.normal(0, 1, ws).astype(np.float32)
    pb = rng.normal(0, 1, bs).astype(np.float32)
    out.update(up_w0=wv.copy(), up_b0=bv.copy(), up_gw=gwv, up_gb=gbv, up_pw0=pw.copy(),
               up_pb0=pb.copy(), up_hyper=np.array([0.9, 0.001, 0.01, 7], np.float64))
    ref.update_params(wv, bv, gwv, gbv, pw, pb, 0.9, 0.001, 0.01, 7)
    out.update(up_w=wv, up_b=bv, up_pw=pw, up_pb=pb)
    np.savez_compressed(os.path.join(HERE, "ref_kernels_small.npz"), **out)


def patches(rng, n, w, h):
    gt = rng.uniform(0, 1, (n, h, w)).astype(np.float32)
    x = np.clip(gt + rng.normal(0, 0.05, gt.shape), 0, 1).astype(np.float32)
    x -= x.mean(axis=(1, 2), keepdims=True)
    return x.astype(np.float32), gt


def reference_train_chain(ref, name, cfg, n_samples, w, h, chunk, epochs, full):
    n1, n2, f1, f2, f3 = cfg
    rng = np.random.default_rng(7 + n1)
    params = make_params(rng, n1, n2, f1, f2, f3, sd=0.1 if f1 == 3 else 0.02)
    x, gt = patches(rng, n_samples, w, h)
    net = NetState(n1, n2, f1, f2, f3, params)
    hyper = dict(momentum=0.9, decay=0.001, lr=np.array([1e-3, 1e-3, 1e-4], np.float32))
    out = dict(cfg=np.array(cfg), dims=np.array([n_samples, w, h, chunk, epochs]),
               x=x, gt=gt, momentum=np.float32(0.9), decay=np.float32(0.001), lr=hyper["lr"])
    for kname, v in params.items():
        out["p0_" + kname] = v
    if full:  # intermediates of the very first chunk
        o1, o2, o3 = ref.net_forward(net, x[:chunk], w, h, chunk)
        d1, d2, d3 = ref.net_backward(net, x[:chunk], gt[:chunk], w, h, chunk, o1, o2, o3)
        out.update(o1=o1, o2=o2, o3=o3, d1=d1, d2=d2, d3=d3)
        for i in range(3):
            out["g1_w%d" % (i + 1)] = net.gw[i].copy()
            out["g1_b%d" % (i + 1)] = net.gb[i].copy()
            net.gw[i][:] = 0
            net.gb[i][:] = 0
    for e in range(epochs):
        ref.net_train_epoch(net, x, gt, w, h, chunk, hyper["momentum"], hyper["decay"],
                            hyper["lr"], True)
        for i in range(3):
            out["e%d_w%d" % (e + 1, i + 1)] = net.w[i].copy()
            out["e%d_b%d" % (e + 1, i + 1)] = net.b[i].copy()
    o1, o2, o3 = ref.net_forward(net, x, w, h, n_samples)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    out["final_sse"] = np.float64(ref.squared_error(gt, o3, w, h, w3, h3, n_samples))
    out["final_o3"] = o3
    np.savez_compressed(os.path.join(HERE, name), **out)


def reference_train_big(ref, name, cfg, n_samples, chunk, epochs, seed, lr):
    """Per-epoch parameters of the reference kernels on a BASELINE-sized training run.  Inputs
    are NOT stored: tests regenerate them from `seed` with tests/helpers.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    n1, n2, f1, f2, f3 = cfg
    rng = np.random.default_rng(seed)
    params = helpers.make_params(rng, n1, n2, f1, f2, f3)
    x, gt = helpers.patches(rng, n_samples, 33, 33)
    net = NetState(n1, n2, f1, f2, f3, params)
    ref.set_num_threads(len(os.sched_getaffinity(0)))
    out = dict(cfg=np.array(cfg), dims=np.array([n_samples, 33, 33, chunk, epochs]),
               seed=np.int64(seed), momentum=np.float32(0.9), decay=np.float32(0.001),
               lr=np.asarray(lr, np.float32),
               x_checksum=np.float64(x.astype(np.float64).sum()),
               gt_checksum=np.float64(gt.astype(np.float64).sum()))
    for e in range(epochs):
        ref.net_train_epoch(net, x, gt, 33, 33, chunk, 0.9, 0.001, out["lr"], True)
        for i in range(3):
            out["e%d_w%d" % (e + 1, i + 1)] = net.w[i].copy()
            out["e%d_b%d" % (e + 1, i + 1)] = net.b[i].copy()
    np.savez_compressed(os.path.join(HERE, name), **out)


def luma_fixtures():
    """Decode the reference's image fixtures with the reference's own stb_image."""
    import subprocess
    import tempfile
    src = r"""
#define STB_IMAGE_IMPLEMENTATION
#include "stb/stb_image.h"
#include <stdio.h>
int main(int argc, char** argv) {
  int w, h, n, want = argv[3][0] - '0';
  unsigned char* d = stbi_load(argv[1], &w, &h, &n, want);
  if (!d) return 1;
  FILE* f = fopen(argv[2], "wb");
  fwrite(d, 1, (size_t)w * h * want, f);
  fclose(f);
  printf("%d %d\n", w, h);
  return 0;
}
"""
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "dec.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "dec")
        subprocess.check_call(["gcc", "-O1", "-w", "-I", os.path.join(REF, "libs/include"), "-o", exe,
                               c, "-lm"])
        for name, out, ch, dims in (
                ("color_grid.png", "color_grid_5x5.rgba", 4, "5 5"),
                ("color_grid2.jpg", "color_grid2_32x32.rgba", 4, "32 32"),
                ("color_grid2_luma_swapped.png", "color_grid2_luma_swapped_32x32.rgb", 3, "32 32")):
            got = subprocess.check_output([exe, os.path.join(REF, "test/data", name),
                                           os.path.join(HERE, out), str(ch)]).decode().strip()
            assert got == dims, (name, got)
    spec = open(os.path.join(REF, "test/specs/ExtractLumaTest.cpp")).read()
    with open(os.path.join(HERE, "luma_goldens.json"), "w") as fh:
        json.dump({"source": "test/specs/ExtractLumaTest.cpp:24-28; SwapLumaTest.cpp:21-24,47-53",
                   "extract_luma_normalized_5x5": c_array(spec, "output"),
                   "margin": 0.005, "swap_padding": 10,
                   "swap_new_luma": "new_luma[i] = i * 1.0f / (luma_w * luma_w), luma_w = 32 - 20"},
                  fh, indent=1)


def main():
    transcribe_reference_specs()
    luma_fixtures()
    copy_config_fixtures()
    ref = Oracle("reference")
    reference_kernel_outputs(ref)
    reference_train_chain(ref, "ref_train_chain.npz", (4, 3, 3, 1, 3), 5, 9, 8, 2, 2, True)
    reference_train_chain(ref, "ref_train_915.npz", (8, 4, 9, 1, 5), 6, 33, 33, 4, 2, False)
    reference_train_big(ref, "ref_train_c2_epoch.npz", (64, 32, 9, 1, 5), 4096, 2048, 2, 4242,
                        [1e-4, 1e-4, 1e-5])
    reference_train_big(ref, "ref_train_c4_epoch.npz", (64, 32, 9, 5, 5), 512, 256, 2, 4343,
                        [1e-4, 1e-4, 1e-5])
    for f in sorted(os.listdir(HERE)):
        print("%8d  %s" % (os.path.getsize(os.path.join(HERE, f)), f))


if __name__ == "__main__":
    main()
