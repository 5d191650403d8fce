"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the oracle
on the same seeded inputs and against the committed golden fixtures.  Run on the B200 box:
    python -m pytest tests -m gpu
Tolerances (north_star): per-pixel max-abs <= 1e-4 on luma in [0,1]; PSNR within 0.01 dB;
per-epoch weights <= 1e-4 relative.  Intermediate FP32 tensors: rtol 1e-4 + atol 1e-5 (the
summation order differs from the oracle's).  8-bit luma output: bit-exact.
"""
import json
import os

import numpy as np
import pytest

import _pkg
from conftest import GOLDEN, load_npz
from helpers import luma_image, make_params, patches, psnr
from oracle.loader import NetState

pytestmark = pytest.mark.gpu
pkg = _pkg.load()

# A/B switches of the library (README "Environment switches"): the whole file can be run under
# any of them as a check of the fallback kernels; the few tests that assert WHICH kernel ran are
# skipped then.
SWITCHED = [k for k in ("SRCNN_FUSED_IMPL", "SRCNN_C5_IMPL", "SRCNN_B3_IMPL", "SRCNN_GW_IMPL",
                        "SRCNN_D1_IMPL") if os.environ.get(k)]
default_paths_only = pytest.mark.skipif(bool(SWITCHED), reason="asserts the default kernel "
                                        "selection; %s is set" % ", ".join(SWITCHED))

RTOL, ATOL = 1e-4, 1e-5


@pytest.fixture(scope="module")
def ctx():
    c = pkg.Context(0)
    yield c
    c.close()


def _json(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def gpu_forward(ctx, x, W, B, k, n, f, skip, w, h, S):
    ow, oh = w - f + 1, h - f + 1
    mi, mW, mB = ctx.upload(x), ctx.upload(W), ctx.upload(B)
    mo = ctx.alloc(4 * S * ow * oh * n)
    ctx.forward_layer(mi, mo, mW, mB, k, n, f, skip, w, h, S)
    out = ctx.read(mo, (S, oh, ow, n))
    for m in (mi, mW, mB, mo):
        ctx.release(m)
    return out


def gpu_deltas(ctx, dn, lo, W, nc, fn, nn, ow, oh, S):
    a, b, c = ctx.upload(dn), ctx.upload(lo), ctx.upload(W)
    t = ctx.alloc(4 * S * ow * oh * nc)
    ctx.deltas(a, b, t, c, nc, fn, nn, ow, oh, S)
    out = ctx.read(t, (S, oh, ow, nc))
    for m in (a, b, c, t):
        ctx.release(m)
    return out


def gpu_backprop(ctx, d, li, gw0, gb0, n, k, f, ow, oh, S):
    a, b = ctx.upload(d), ctx.upload(li)
    gw, gb = ctx.upload(gw0), ctx.upload(gb0)
    ctx.backpropagate(a, b, gw, gb, n, k, f, ow, oh, S)
    r = ctx.read(gw, (gw0.size,)), ctx.read(gb, (gb0.size,))
    for m in (a, b, gw, gb):
        ctx.release(m)
    return r


# ------------------------------------------------------------------ reference goldens
def test_forward_reference_layer_cases(ctx):
    """reference: test/specs/LayerTest.cpp:97-130, test/data/test_cases.json"""
    for name, c in _json("layer_cases.json")["cases"].items():
        k, n, f = c["n_prev_filter_cnt"], c["current_filter_count"], c["f_spatial_size"]
        out = gpu_forward(ctx, np.array(c["input"], np.float32), np.array(c["weights"], np.float32),
                          np.array(c["bias"], np.float32), k, n, f, False, c["input_w"],
                          c["input_h"], 1)
        np.testing.assert_allclose(out.reshape(-1), np.array(c["output"], np.float32),
                                   atol=5.1e-4, rtol=0, err_msg=name)


def test_deltas_reference_case(ctx):
    """reference: test/specs/LayerDeltasTest.cpp:33-126"""
    c = _json("layer_deltas_case.json")
    lo = np.maximum(np.array(c["input_x"], np.float32), 0)
    out = gpu_deltas(ctx, np.array(c["deltas"], np.float32), lo, np.array(c["weights"], np.float32),
                     c["n_curr"], c["f_next"], c["n_next"], c["out_w"], c["out_h"], 1)
    np.testing.assert_allclose(out.reshape(-1), np.array(c["expected_output"], np.float32),
                               atol=2e-6, rtol=0)


def test_backpropagate_reference_case(ctx):
    """reference: test/specs/BackpropagationTest.cpp:31-90 (accumulates onto 1.5)"""
    c = _json("backprop_case.json")
    gw0 = np.full(c["f"] ** 2 * c["k"] * c["n"], c["grad_w_init"], np.float32)
    gw, gb = gpu_backprop(ctx, np.array(c["deltas"], np.float32), np.array(c["input"], np.float32),
                          gw0, np.zeros(c["n"], np.float32), c["n"], c["k"], c["f"], c["out_w"],
                          c["out_h"], 1)
    np.testing.assert_allclose(gw, np.array(c["expected_weights"], np.float32), atol=6e-5, rtol=0)
    np.testing.assert_allclose(gb, np.array(c["expected_bias"], np.float32), atol=1e-6, rtol=0)


def test_backpropagate_big_data_does_not_crash(ctx):
    """reference: BackpropagationTest.cpp data set 2 (k=32 n=16 f=3 on 1024x1024 deltas)"""
    n, k, f, ow, oh = 16, 32, 3, 1024, 1024
    d = ctx.zeros(ow * oh * n)
    li = ctx.zeros((ow + f - 1) * (oh + f - 1) * k)
    gw, gb = ctx.zeros(f * f * k * n), ctx.zeros(n)
    ctx.backpropagate(d, li, gw, gb, n, k, f, ow, oh, 1)
    ctx.block()
    assert not ctx.read(gw, (f * f * k * n,)).any()
    for m in (d, li, gw, gb):
        ctx.release(m)


def test_committed_reference_kernel_outputs(ctx):
    """outputs of the reference's own kernels (tests/golden/ref_kernels_small.npz)"""
    g = load_npz("ref_kernels_small.npz")
    S, k, n, f, w, h = (int(v) for v in g["fw_shape"])
    np.testing.assert_allclose(gpu_forward(ctx, g["fw_x"], g["fw_W"], g["fw_B"], k, n, f, False, w, h, S),
                               g["fw_relu"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(gpu_forward(ctx, g["fw_x"], g["fw_W"], g["fw_B"], k, n, f, True, w, h, S),
                               g["fw_lin"], rtol=RTOL, atol=ATOL)
    S, nc, fn, nn, ow, oh = (int(v) for v in g["dl_shape"])
    np.testing.assert_allclose(gpu_deltas(ctx, g["dl_next"], g["dl_out"], g["dl_W"], nc, fn, nn, ow, oh, S),
                               g["dl_result"], rtol=RTOL, atol=ATOL)
    S, k, n, f, ow, oh = (int(v) for v in g["bp_shape"])
    gw, gb = gpu_backprop(ctx, g["bp_d"], g["bp_in"], g["bp_gw0"], g["bp_gb0"], n, k, f, ow, oh, S)
    np.testing.assert_allclose(gw, g["bp_gw"], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(gb, g["bp_gb"], rtol=RTOL, atol=1e-4)
    # last layer delta + squared error
    S, ah, aw = g["ll_algo"].shape
    gh, gwid = g["ll_gt"].shape[1:]
    mg, ma = ctx.upload(g["ll_gt"]), ctx.upload(g["ll_algo"])
    mt, ms = ctx.alloc(4 * S * ah * aw), ctx.alloc(4)
    ctx.last_layer_delta(mg, ma, mt, gwid, gh, aw, ah, S)
    np.testing.assert_array_equal(ctx.read(mt, (S, ah, aw)), g["ll_delta"])
    ctx.squared_error(mg, ma, ms, gwid, gh, aw, ah, S)
    assert float(ctx.read(ms, (1,))[0]) == pytest.approx(float(g["ll_sse"]), rel=1e-6)
    # update with weight decay
    m, dec, lr, batch = g["up_hyper"]
    hs = [ctx.upload(g[k_]) for k_ in ("up_w0", "up_b0", "up_gw", "up_gb", "up_pw0", "up_pb0")]
    ctx.update_params(*hs, float(m), float(dec), float(lr), int(batch), g["up_w0"].size, g["up_b0"].size)
    for hnd, key in zip((hs[0], hs[1], hs[4], hs[5]), ("up_w", "up_b", "up_pw", "up_pb")):
        np.testing.assert_allclose(ctx.read(hnd, g[key].shape), g[key], rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ per-kernel vs oracle
FWD_SHAPES = [
    # k, n, f, w, h, S
    (1, 64, 9, 33, 33, 3), (64, 32, 1, 25, 25, 3), (32, 1, 5, 25, 25, 3), (64, 32, 5, 25, 25, 2),
    (1, 128, 9, 40, 21, 1), (128, 64, 1, 17, 30, 2), (64, 1, 5, 31, 18, 2),
    (1, 32, 9, 20, 20, 1), (32, 16, 1, 12, 12, 2), (16, 1, 5, 12, 12, 2),
    (5, 7, 3, 8, 6, 3), (3, 70, 1, 5, 5, 1), (2, 3, 5, 5, 5, 1), (1, 1, 1, 1, 1, 1),
]


@pytest.mark.parametrize("shape", FWD_SHAPES)
def test_forward_vs_oracle(ctx, port, shape):
    k, n, f, w, h, S = shape
    rng = np.random.default_rng(hash(shape) % 2**32)
    x = rng.normal(0, 1, (S, h, w, k)).astype(np.float32)
    W = rng.normal(0, 1.0 / np.sqrt(f * f * k), f * f * k * n).astype(np.float32)
    B = rng.normal(0, 0.1, n).astype(np.float32)
    for skip in (False, True):
        np.testing.assert_allclose(gpu_forward(ctx, x, W, B, k, n, f, skip, w, h, S),
                                   port.forward(x, W, B, k, n, f, skip, w, h, S),
                                   rtol=RTOL, atol=ATOL)


DELTA_SHAPES = [
    # n_curr, f_next, n_next, out_w, out_h, S
    (32, 5, 1, 25, 25, 3), (64, 1, 32, 25, 25, 3), (64, 5, 32, 25, 25, 2), (64, 5, 1, 21, 30, 2),
    (128, 1, 64, 9, 11, 2), (16, 5, 1, 12, 12, 2), (32, 1, 16, 12, 12, 2),
    (4, 3, 5, 6, 5, 2), (3, 3, 2, 3, 3, 1), (70, 1, 3, 4, 4, 1),
]


@pytest.mark.parametrize("shape", DELTA_SHAPES)
def test_deltas_vs_oracle(ctx, port, shape):
    nc, fn, nn, ow, oh, S = shape
    rng = np.random.default_rng(hash(shape) % 2**32)
    dn = rng.normal(0, 1, (S, oh - fn + 1, ow - fn + 1, nn)).astype(np.float32)
    lo = np.maximum(rng.normal(0, 1, (S, oh, ow, nc)), 0).astype(np.float32)
    W = rng.normal(0, 0.3, fn * fn * nc * nn).astype(np.float32)
    np.testing.assert_allclose(gpu_deltas(ctx, dn, lo, W, nc, fn, nn, ow, oh, S),
                               port.deltas(dn, lo, W, nc, fn, nn, ow, oh, S), rtol=RTOL, atol=ATOL)


@default_paths_only
@pytest.mark.parametrize("w1,h1,S", [(25, 25, 170), (30, 19, 150), (25, 25, 333),
                                     # maps lower than the 8-row accumulator ring / the 5 filter rows
                                     (9, 5, 460), (8, 7, 520), (12, 9, 350), (40, 13, 110)])
def test_conv5_tensor_core_kernels_vs_oracle(ctx, port, w1, h1, S):
    """The 9-5-5 network's layer 2 on the tensor cores (conv5_tc.cuh): forward 64 -> 32 and the
    layer-1 deltas, as virtual-image implicit GEMMs with FP16-split operands, at sample counts
    that take that path (>= 32 strips of 128 columns).  reference:
    src/kernel/layer_uber_kernel.cl:36-96, src/kernel/layer_deltas.cl:42-127."""
    rng = np.random.default_rng(1000 * w1 + S)
    port.set_num_threads(len(os.sched_getaffinity(0)))
    k, n, f = 64, 32, 5
    x = np.maximum(rng.normal(0, 1, (S, h1, w1, k)), 0).astype(np.float32)       # an out1: >= 0
    W = rng.normal(0, 1.0 / np.sqrt(f * f * k), f * f * k * n).astype(np.float32)
    B = rng.normal(0, 0.1, n).astype(np.float32)
    n0 = ctx.launch_count()
    got = gpu_forward(ctx, x, W, B, k, n, f, False, w1, h1, S)
    assert ctx.launch_count() - n0 >= 2, "the tensor-core path (absmax + contraction) did not run"
    np.testing.assert_allclose(got, port.forward(x, W, B, k, n, f, False, w1, h1, S),
                               rtol=RTOL, atol=ATOL)
    # deltas of layer 1 from the deltas of layer 2 (tiny values, like real ones)
    dn = (rng.normal(0, 1, (S, h1 - 4, w1 - 4, n)) * 3e-4).astype(np.float32)
    n0 = ctx.launch_count()
    got = gpu_deltas(ctx, dn, x, W, k, f, n, w1, h1, S)
    assert ctx.launch_count() - n0 >= 2, "the tensor-core path (absmax + contraction) did not run"
    exp = port.deltas(dn, x, W, k, f, n, w1, h1, S)
    np.testing.assert_allclose(got, exp, rtol=RTOL, atol=ATOL * 3e-4)
    assert np.abs(exp).max() > 1e-5


BP_SHAPES = [
    # n, k, f, out_w, out_h, S
    (64, 1, 9, 25, 25, 5), (32, 64, 1, 25, 25, 5), (1, 32, 5, 21, 21, 5), (32, 64, 5, 21, 21, 2),
    (128, 1, 9, 12, 9, 2), (64, 128, 1, 10, 10, 2), (16, 32, 1, 8, 8, 3),
    (4, 3, 3, 5, 4, 3), (3, 2, 3, 3, 3, 1), (1, 1, 1, 1, 1, 1), (70, 3, 1, 9, 2, 2),
    # the 5x5 layer 2 of 9-5-5 at sample counts that take the MN-major tensor-core kernel
    # (wgrad5_tc.cuh)
    (32, 64, 5, 21, 21, 170), (32, 64, 5, 26, 15, 150),
]


@pytest.mark.parametrize("shape", BP_SHAPES)
def test_backpropagate_vs_oracle(ctx, port, shape):
    n, k, f, ow, oh, S = shape
    rng = np.random.default_rng(hash(shape) % 2**32)
    port.set_num_threads(len(os.sched_getaffinity(0)))
    d = rng.normal(0, 1, (S, oh, ow, n)).astype(np.float32)
    li = rng.normal(0, 1, (S, oh + f - 1, ow + f - 1, k)).astype(np.float32)
    gw0 = rng.normal(0, 1, f * f * k * n).astype(np.float32)
    gb0 = rng.normal(0, 1, n).astype(np.float32)
    egw, egb = gw0.copy(), gb0.copy()
    port.backpropagate(d, li, egw, egb, n, k, f, ow, oh, S)
    gw, gb = gpu_backprop(ctx, d, li, gw0, gb0, n, k, f, ow, oh, S)
    scale = np.sqrt(S * ow * oh)   # |sum of P unit-variance products| ~ sqrt(P)
    np.testing.assert_allclose(gw, egw, rtol=RTOL, atol=2e-5 * scale)
    np.testing.assert_allclose(gb, egb, rtol=RTOL, atol=2e-5 * scale)


def test_backpropagate_is_deterministic(ctx):
    """The reference races on grad_w across samples (backpropagate.cl:110); ours must give the
    same bits on every run."""
    n, k, f, ow, oh, S = 64, 1, 9, 25, 25, 16
    rng = np.random.default_rng(5)
    d = rng.normal(0, 1, (S, oh, ow, n)).astype(np.float32)
    li = rng.normal(0, 1, (S, oh + f - 1, ow + f - 1, k)).astype(np.float32)
    z = np.zeros(f * f * k * n, np.float32), np.zeros(n, np.float32)
    a = gpu_backprop(ctx, d, li, z[0], z[1], n, k, f, ow, oh, S)
    b = gpu_backprop(ctx, d, li, z[0], z[1], n, k, f, ow, oh, S)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_elementwise_and_reductions_vs_oracle(ctx, port):
    rng = np.random.default_rng(8)
    # reference: LastLayerDeltaTest.cpp (6x6 in 14x14, poisoned border), S>1 added
    S, aw, ah, pad = 3, 6, 6, 4
    gt = np.full((S, ah + 2 * pad, aw + 2 * pad), 99999.0, np.float32)
    gt[:, pad:pad + ah, pad:pad + aw] = rng.uniform(0, 2.56, (S, ah, aw))
    algo = np.maximum(rng.uniform(-1, 1.56, (S, ah, aw)), 0).astype(np.float32)
    mg, ma, mt, ms = ctx.upload(gt), ctx.upload(algo), ctx.alloc(4 * S * aw * ah), ctx.alloc(4)
    ctx.last_layer_delta(mg, ma, mt, aw + 2 * pad, ah + 2 * pad, aw, ah, S)
    np.testing.assert_array_equal(ctx.read(mt, (S, ah, aw)),
                                  port.last_layer_delta(gt, algo, aw + 2 * pad, ah + 2 * pad, aw, ah, S))
    ctx.squared_error(mg, ma, ms, aw + 2 * pad, ah + 2 * pad, aw, ah, S)
    assert float(ctx.read(ms, (1,))[0]) == pytest.approx(
        port.squared_error(gt, algo, aw + 2 * pad, ah + 2 * pad, aw, ah, S), rel=1e-6)
    # reference: SquaredErrorTest.cpp at its full size 1000x2000 in 1008x2008
    aw, ah, pad = 1000, 2000, 4
    gt = np.full((ah + 2 * pad, aw + 2 * pad), 99999.0, np.float32)
    gt[pad:pad + ah, pad:pad + aw] = rng.integers(0, 256, (ah, aw))
    algo = (rng.integers(0, 2560, (ah, aw)) / 10.0).astype(np.float32)
    mg2, ma2 = ctx.upload(gt), ctx.upload(algo)
    ctx.squared_error(mg2, ma2, ms, aw + 2 * pad, ah + 2 * pad, aw, ah, 1)
    assert float(ctx.read(ms, (1,))[0]) == pytest.approx(
        port.squared_error(gt, algo, aw + 2 * pad, ah + 2 * pad, aw, ah, 1), rel=1e-6)
    # reference: SumTest.cpp / SubtractFromAllTest.cpp (0..899)
    data = np.arange(900, dtype=np.float32)
    md = ctx.upload(data)
    ctx.sum(md, 900, False, ms)
    assert float(ctx.read(ms, (1,))[0]) == 404550.0
    ctx.sum(md, 900, True, ms)
    assert float(ctx.read(ms, (1,))[0]) == pytest.approx(port.sum(data, True), rel=1e-7)
    ctx.sub_from_all(md, 450.0, 900)
    np.testing.assert_array_equal(ctx.read(md, (900,)), data - 450.0)
    # a length that is not a multiple of anything, > one reduction pass
    big = rng.normal(0, 1, 1_234_567).astype(np.float32)
    mb = ctx.upload(big)
    ctx.sum(mb, big.size, False, ms)
    assert float(ctx.read(ms, (1,))[0]) == pytest.approx(port.sum(big), rel=1e-5, abs=1e-2)


def test_update_parameters_vs_oracle(ctx, port):
    """reference: UpdateParametersTest.cpp (2*400*5^2 weights, 400 biases, momentum .8,
    lr .001, batch 2) plus non-zero weight decay."""
    rng = np.random.default_rng(13)
    ws, bs = 2 * 400 * 25, 400
    for decay in (0.0, 0.05):
        arrs = dict(w=rng.integers(0, 2560, ws) / 10.0, b=rng.integers(0, 2560, bs) / 10.0,
                    gw=rng.integers(0, 2560, ws) / 100.0, gb=rng.integers(0, 2560, bs) / 100.0,
                    pw=rng.integers(0, 2560, ws) / 10.0, pb=rng.integers(0, 2560, bs) / 10.0)
        arrs = {k: v.astype(np.float32) for k, v in arrs.items()}
        hs = {k: ctx.upload(v) for k, v in arrs.items()}
        ctx.update_params(hs["w"], hs["b"], hs["gw"], hs["gb"], hs["pw"], hs["pb"], 0.8, decay,
                          0.001, 2, ws, bs)
        e = {k: v.copy() for k, v in arrs.items()}
        port.update_params(e["w"], e["b"], e["gw"], e["gb"], e["pw"], e["pb"], 0.8, decay, 0.001, 2)
        for k in ("w", "b", "pw", "pb"):
            np.testing.assert_allclose(ctx.read(hs[k], arrs[k].shape), e[k], rtol=1e-6, atol=1e-5)


def test_luma_kernels_bit_exact(ctx, port):
    """extract_luma / swap_luma (reference: ExtractLumaTest.cpp, SwapLumaTest.cpp): float luma
    to 1 ulp-ish, 8-bit output bit-exact against the oracle."""
    rng = np.random.default_rng(17)
    h, w, pad = 37, 53, 6
    rgba = rng.integers(0, 256, (h, w, 4)).astype(np.uint8)
    mi = ctx.upload(rgba, np.uint8)
    ml = ctx.alloc(4 * w * h)
    for norm in (True, False):
        ctx.extract_luma(mi, ml, w, h, norm)
        np.testing.assert_array_equal(ctx.read(ml, (h, w)), port.extract_luma(rgba, norm))
    new_luma = rng.uniform(-0.1, 1.1, (h - 2 * pad, w - 2 * pad)).astype(np.float32)
    mn = ctx.upload(new_luma)
    mt = ctx.alloc(w * h * 3)
    ctx.swap_luma(mi, mn, mt, w, h, w - 2 * pad, h - 2 * pad)
    np.testing.assert_array_equal(ctx.read(mt, (h, w, 3), np.uint8),
                                  port.swap_luma(rgba, new_luma, w - 2 * pad, h - 2 * pad))


def test_luma_kernels_reference_fixtures(ctx):
    """The reference's own luma fixtures (decoded by its stb_image, tests/golden/make_golden.py):
    all 25 goldens of ExtractLumaTest.cpp:24-28 (both data sets) and the exact image of
    SwapLumaTest.cpp:39-90."""
    rd = lambda n, shape: np.fromfile(os.path.join(GOLDEN, n), np.uint8).reshape(shape)
    gold = _json("luma_goldens.json")
    grid = rd("color_grid_5x5.rgba", (5, 5, 4))
    exp = np.array(gold["extract_luma_normalized_5x5"], np.float32).reshape(5, 5)
    mi, ml = ctx.upload(grid, np.uint8), ctx.alloc(4 * 25)
    ctx.extract_luma(mi, ml, 5, 5, True)
    np.testing.assert_allclose(ctx.read(ml, (5, 5)), exp, atol=gold["margin"], rtol=0)
    ctx.extract_luma(mi, ml, 5, 5, False)
    np.testing.assert_allclose(ctx.read(ml, (5, 5)), exp * 255.0, atol=255 * gold["margin"], rtol=0)
    img = rd("color_grid2_32x32.rgba", (32, 32, 4))
    expected = rd("color_grid2_luma_swapped_32x32.rgb", (32, 32, 3))
    lw = 32 - 2 * gold["swap_padding"]
    new_luma = (np.arange(lw * lw, dtype=np.float32) * np.float32(1.0) / np.float32(lw * lw))
    mi, mn, mt = ctx.upload(img, np.uint8), ctx.upload(new_luma), ctx.alloc(32 * 32 * 3)
    ctx.swap_luma(mi, mn, mt, 32, 32, lw, lw)
    np.testing.assert_array_equal(ctx.read(mt, (32, 32, 3), np.uint8), expected)


# ------------------------------------------------------------------ errors
def test_error_behaviour(ctx):
    """Validation mirrors the reference's (src/DataPipeline.cpp:339-356, Context.cpp:235-341):
    failures are reported, never silently computed."""
    small = ctx.alloc(16)
    with pytest.raises(pkg.SrcnnError, match="too small"):
        ctx.forward_layer(small, small, small, small, 1, 64, 9, False, 33, 33, 1)
    with pytest.raises(pkg.SrcnnError, match="smaller than filter"):
        ctx.forward_layer(small, small, small, small, 1, 1, 9, False, 5, 5, 1)
    with pytest.raises(pkg.SrcnnError, match="more then is allocated"):
        ctx.read(small, (100,))
    with pytest.raises(pkg.SrcnnError, match="invalid memory handle"):
        ctx.fill_float(pkg.NULL_MEM, 0.0)
    ctx.release(small)
    with pytest.raises(pkg.SrcnnError, match="invalid memory handle"):
        ctx.read(small, (1,))
    a, b = ctx.alloc(64), ctx.alloc(32)
    with pytest.raises(pkg.SrcnnError, match="after dst end"):
        ctx.copy(a, b)


# ------------------------------------------------------------------ whole net
NETS = [
    # n1, n2, f1, f2, f3, w, h, S
    (64, 32, 9, 1, 5, 33, 33, 4), (64, 32, 9, 5, 5, 33, 33, 2), (128, 64, 9, 1, 5, 40, 29, 2),
    (32, 16, 9, 1, 5, 33, 33, 3), (8, 4, 9, 1, 5, 33, 33, 3), (4, 3, 3, 1, 3, 9, 8, 3),
    # the fused tensor-core forward treats a chunk as one virtual image [h][S*w]: ragged
    # patches over several 124-column strips, 1x1 outputs, and tall samples (several row bands)
    (64, 32, 9, 1, 5, 40, 29, 7), (64, 32, 9, 1, 5, 33, 33, 37), (64, 32, 9, 1, 5, 13, 13, 5),
    (64, 32, 9, 1, 5, 20, 150, 3),
    # sample counts that take the tensor-core kernels of the backward pass as bench.py's chunks
    # do: bwd3_tc (>= 32 strips of out2 columns), and for 9-5-5 the layer-1-only forward,
    # conv5_tc (forward, layer-1 deltas) and wgrad5_tc
    (64, 32, 9, 1, 5, 33, 33, 170), (64, 32, 9, 5, 5, 33, 33, 200), (64, 32, 9, 1, 5, 40, 29, 140),
    (64, 32, 9, 5, 5, 36, 31, 180),
]


@pytest.mark.parametrize("cfg", NETS)
def test_train_chunk_vs_oracle(ctx, port, cfg):
    """forward + deltas + gradients of one chunk: every intermediate buffer and the six
    gradient tensors against the oracle."""
    n1, n2, f1, f2, f3, w, h, S = cfg
    rng = np.random.default_rng(sum(cfg))
    port.set_num_threads(len(os.sched_getaffinity(0)))
    params = make_params(rng, n1, n2, f1, f2, f3)
    x, gt = patches(rng, S, w, h)
    on = NetState(n1, n2, f1, f2, f3, params)
    o1, o2, o3 = port.net_forward(on, x, w, h, S)

    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    mi, mg = ctx.upload(x), ctx.upload(gt)
    work = ctx.alloc(net.train_workspace_bytes(w, h, S))
    net.train_chunk(mi, mg, w, h, S, work)
    flat = ctx.read(work, (net.train_workspace_bytes(w, h, S) // 4,))
    off = 0   # sub-buffers are 256-byte aligned inside the workspace (srcnn_train_chunk)
    got = {}
    for name, exp in (("out1", o1), ("out2", o2), ("out3", o3), ("d1", o1), ("d2", o2), ("d3", o3)):
        got[name] = flat[off:off + exp.size].reshape(exp.shape)
        off += ((exp.size * 4 + 255) // 256 * 256) // 4
    for name, exp in (("out1", o1), ("out2", o2), ("out3", o3)):
        np.testing.assert_allclose(got[name], exp, rtol=RTOL, atol=ATOL, err_msg=name)
    # The backward pass switches on [activation > 0] (ReLU derivative, and quirk Q2 on the
    # last layer): an activation within rounding noise of zero may switch differently on the
    # two sides, which changes a delta by its full value.  The backward step is therefore
    # checked against the oracle's backward pass run on the activations the GPU produced
    # (already shown equal to the oracle's forward pass to 1e-5 above).
    d1, d2, d3 = port.net_backward(on, x, gt, w, h, S, np.ascontiguousarray(got["out1"]),
                                   np.ascontiguousarray(got["out2"]),
                                   np.ascontiguousarray(got["out3"]))
    for name, exp in (("d1", d1), ("d2", d2), ("d3", d3)):
        if name == "d1" and not net.train_materializes_d1():
            continue   # lives inside the layer-1 gradient kernel: checked through gw1 / gb1 below
        np.testing.assert_allclose(got[name], exp, rtol=RTOL, atol=ATOL, err_msg=name)
    g = net.grads()
    for l in range(3):
        # sums over S * pixels terms: the absolute floor follows the size of the tensor's entries
        # (a near-cancelling entry of a gradient whose other entries are ~500 is good to ~1e-6 of
        # those, not to 1e-4 absolute)
        for key, exp in (("w%d" % (l + 1), on.gw[l]), ("b%d" % (l + 1), on.gb[l])):
            np.testing.assert_allclose(g[key], exp, rtol=RTOL, atol=1e-4 + 2e-6 * float(np.abs(exp).max()),
                                       err_msg="g" + key)


@default_paths_only
def test_layer1_deltas_fused_and_separate_agree(ctx, port):
    """The layer-1 deltas normally live inside the layer-1 gradient kernel; with
    SRCNN_D1_IMPL=separate they are materialised by their own launch.  Same gradients either
    way, several tiles per CTA with a ragged last tile, and d1 against the oracle where it is
    written."""
    n1, n2, f1, f2, f3, w, h, S = 64, 32, 9, 1, 5, 33, 33, 301
    rng = np.random.default_rng(301)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x, gt = patches(rng, S, w, h)
    os.environ["SRCNN_D1_IMPL"] = "separate"
    try:
        ctx2 = pkg.Context(0)
    finally:
        del os.environ["SRCNN_D1_IMPL"]
    grads = []
    for c in (ctx, ctx2):
        net = pkg.Net(c, n1, n2, f1, f2, f3, params)
        assert net.train_materializes_d1() == (c is ctx2)
        work = c.alloc(net.train_workspace_bytes(w, h, S))
        net.train_chunk(c.upload(x), c.upload(gt), w, h, S, work)
        grads.append(net.grads())
        c.release(work)
    for k in grads[0]:
        scale = float(np.abs(grads[1][k]).max())
        assert scale > 0
        assert float(np.abs(grads[0][k] - grads[1][k]).max()) <= 2e-5 * scale, k
    ctx2.close()


def test_train_chunks_from_host_buffers(ctx, port):
    """execute_batch(backpropagate=true) on host-resident samples (uploads pipelined with the
    training of the previous chunk, ragged last chunk): gradients equal those of one device-
    resident chunk over the same samples."""
    n1, n2, f1, f2, f3, w, h, n = 64, 32, 9, 1, 5, 33, 33, 7
    rng = np.random.default_rng(707)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x, gt = patches(rng, n, w, h)
    net_a = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    work = ctx.alloc(net_a.train_workspace_bytes(w, h, n))
    net_a.train_chunk(ctx.upload(x), ctx.upload(gt), w, h, n, work)
    ga = net_a.grads()
    net_b = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    hx, hg = pkg.PinnedBuffer(x.shape), pkg.PinnedBuffer(gt.shape)
    hx.array[:], hg.array[:] = x, gt
    net_b.train_chunks_host(hx.array, hg.array, w, h, 3, work)
    ctx.block()
    gb = net_b.grads()
    for k in ga:
        np.testing.assert_allclose(gb[k], ga[k], rtol=RTOL, atol=1e-5, err_msg=k)
        assert np.abs(ga[k]).max() > 0


@pytest.mark.parametrize("name", ["ref_train_chain.npz", "ref_train_915.npz"])
def test_training_epochs_vs_committed_reference(ctx, name):
    """2 epochs x 2 chunks, momentum + weight decay: per-epoch parameters within 1e-4 relative
    of what the reference's own kernels produce (fixture generated by make_golden.py)."""
    g = load_npz(name)
    n1, n2, f1, f2, f3 = (int(v) for v in g["cfg"])
    ns, w, h, chunk, epochs = (int(v) for v in g["dims"])
    params = {k[3:]: g[k] for k in g if k.startswith("p0_")}
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    x, gt = g["x"], g["gt"]
    work = ctx.alloc(net.train_workspace_bytes(w, h, chunk))
    for e in range(epochs):
        for i in range(0, ns, chunk):
            S = min(chunk, ns - i)
            mi, mg = ctx.upload(x[i:i + S]), ctx.upload(gt[i:i + S])
            net.train_chunk(mi, mg, w, h, S, work)
        net.update_all(ns, float(g["momentum"]), float(g["decay"]), g["lr"])
        p = net.params()
        for l in range(3):
            np.testing.assert_allclose(p["w%d" % (l + 1)], g["e%d_w%d" % (e + 1, l + 1)],
                                       rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(p["b%d" % (l + 1)], g["e%d_b%d" % (e + 1, l + 1)],
                                       rtol=1e-4, atol=1e-7)
        assert not any(v.any() for v in net.grads().values()), "accumulators must be zeroed"
    # validation SSE after training
    mi, mg, tgt = ctx.upload(x), ctx.upload(gt), ctx.alloc(4)
    work2 = ctx.alloc(net.train_workspace_bytes(w, h, ns))
    net.validate_chunk(mi, mg, w, h, ns, work2, tgt)
    assert float(ctx.read(tgt, (1,))[0]) == pytest.approx(float(g["final_sse"]), rel=1e-4)


def _epoch_fixture_inputs(g):
    from helpers import make_params, patches
    n1, n2, f1, f2, f3 = (int(v) for v in g["cfg"])
    ns, w, h, chunk, epochs = (int(v) for v in g["dims"])
    rng = np.random.default_rng(int(g["seed"]))
    params = make_params(rng, n1, n2, f1, f2, f3)
    x, gt = patches(rng, ns, w, h)
    assert float(x.astype(np.float64).sum()) == pytest.approx(float(g["x_checksum"]), rel=1e-12)
    return (n1, n2, f1, f2, f3), (ns, w, h, chunk, epochs), params, x, gt


def _assert_params_close(p, g, e, what):
    for l in range(3):
        np.testing.assert_allclose(p["w%d" % (l + 1)], g["e%d_w%d" % (e, l + 1)], rtol=1e-4,
                                   atol=1e-7, err_msg="%s epoch %d w%d" % (what, e, l + 1))
        np.testing.assert_allclose(p["b%d" % (l + 1)], g["e%d_b%d" % (e, l + 1)], rtol=1e-4,
                                   atol=1e-7, err_msg="%s epoch %d b%d" % (what, e, l + 1))


@pytest.mark.parametrize("name,route", [("ref_train_c2_epoch.npz", "device"),
                                        ("ref_train_c2_epoch.npz", "host"),
                                        ("ref_train_c4_epoch.npz", "device")])
def test_baseline_sized_epochs_vs_committed_reference(ctx, name, route):
    """The BENCHMARKED training configuration, exactly as bench.py runs it: config C2 (9-1-5
    n1=64 n2=32, 4096 patches 33x33, chunks of 2048 -> the persistent tensor-core kernels
    forward_fused_hp<BATCH>, bwd3_fused, wgrad1_fused_tc, wgrad2_tc), momentum 0.9, decay 1e-3,
    lr 1e-4/1e-4/1e-5, srcnn_update_all -- and the 9-5-5 network of C4 at chunks of 256.  All six
    parameter tensors after epochs 1 and 2 must be within 1e-4 relative of what the reference's
    own kernels (oracle/_ref, fixture committed by make_golden.py) produce.  `host` goes through
    srcnn_train_chunks_host (pinned host samples, the e2e route of the bench)."""
    g = load_npz(name)
    cfg, (ns, w, h, chunk, epochs), params, x, gt = _epoch_fixture_inputs(g)
    net = pkg.Net(ctx, *cfg, params)
    work = ctx.alloc(net.train_workspace_bytes(w, h, chunk))
    mi, mg = ctx.upload(x), ctx.upload(gt)
    per = 4 * w * h
    views = [(ctx.wrap(ctx.mem_ptr(mi) + i * per, min(chunk, ns - i) * per),
              ctx.wrap(ctx.mem_ptr(mg) + i * per, min(chunk, ns - i) * per), min(chunk, ns - i))
             for i in range(0, ns, chunk)]
    for e in range(1, epochs + 1):
        if route == "host":
            net.train_chunks_host(x, gt, w, h, chunk, work)
        else:
            for vi, vg, S in views:
                net.train_chunk(vi, vg, w, h, S, work)
        net.update_all(ns, float(g["momentum"]), float(g["decay"]), g["lr"])
        _assert_params_close(net.params(), g, e, "%s/%s" % (name, route))
    for m in (work, mi, mg):
        ctx.release(m)


@default_paths_only
@pytest.mark.parametrize("cfg", [(64, 32, 9, 1, 5), (64, 32, 9, 5, 5)])
def test_chunk_views_need_only_float_alignment(ctx, cfg):
    """A chunk that starts at a sample offset which is not a multiple of 4 samples is only
    float-aligned (33*33*4 bytes per sample): it must take the same kernels and give the same bits
    as the 16-byte aligned copy of the same samples."""
    rng = np.random.default_rng(77)
    S, w = 171, 33
    params = make_params(rng, *cfg)
    x, gt = patches(rng, S + 1, w, w)
    per = 4 * w * w
    big_i, big_g = ctx.upload(x), ctx.upload(gt)
    results = []
    for shifted in (False, True):
        net = pkg.Net(ctx, *cfg, params)
        work = ctx.alloc(net.train_workspace_bytes(w, w, S))
        if shifted:   # samples 1..S of the big arrays: a view at +4356 bytes
            mi = ctx.wrap(ctx.mem_ptr(big_i) + per, S * per)
            mg = ctx.wrap(ctx.mem_ptr(big_g) + per, S * per)
        else:
            mi, mg = ctx.upload(x[1:]), ctx.upload(gt[1:])
        n0 = ctx.launch_count()
        net.train_chunk(mi, mg, w, w, S, work)
        results.append((ctx.launch_count() - n0, net.grads()))
        ctx.release(work)
    assert results[0][0] == results[1][0], "the float-aligned view took a different (slower) path"
    for k in results[0][1]:
        np.testing.assert_array_equal(results[0][1][k], results[1][1][k], err_msg=k)


INFER = [
    # n1, n2, f1, f2, f3, w, h, S
    (64, 32, 9, 1, 5, 256, 256, 1),     # BASELINE config C1
    (64, 32, 9, 1, 5, 13, 13, 1),       # 1x1 output
    (64, 32, 9, 1, 5, 141, 77, 2),      # ragged, S>1
    (64, 32, 9, 1, 5, 33, 33, 19),      # validation patches: batch as a virtual wide image
    (64, 32, 9, 1, 5, 500, 40, 3),      # wide and flat, S>1
    (64, 32, 9, 1, 5, 600, 30, 2),      # S>1 but too wide for the virtual-image variant
    (64, 32, 9, 5, 5, 96, 80, 1),       # 9-5-5
    (128, 64, 9, 1, 5, 120, 67, 1),     # C5's network (the wide tensor-core kernel)
    (128, 64, 9, 1, 5, 300, 150, 2),    # ... more strips than one, S>1
    (128, 64, 9, 1, 5, 33, 33, 11),     # ... validation patches as a virtual wide image
    (128, 64, 9, 1, 5, 13, 14, 1),      # ... 1x2 output
    (32, 16, 9, 1, 5, 64, 64, 1),       # example_config.json
    (4, 3, 3, 1, 3, 20, 11, 2),         # no fused instantiation -> three-launch path
]


@pytest.mark.parametrize("cfg", INFER)
def test_inference_vs_oracle(ctx, port, cfg):
    """Whole forward pass (the fused launch where instantiated): max-abs <= 1e-4 on luma,
    PSNR (from the squared-error sum) within 0.01 dB."""
    n1, n2, f1, f2, f3, w, h, S = cfg
    rng = np.random.default_rng(sum(cfg) + 1)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x = np.stack([luma_image(rng, h, w) for _ in range(S)])
    gt = np.stack([luma_image(rng, h, w) for _ in range(S)])
    on = NetState(n1, n2, f1, f2, f3, params)
    _, _, e3 = port.net_forward(on, x, w, h, S)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (w1, h1), (w2, h2), (w3, h3) = net.out_dims(w, h)
    mi, mo = ctx.upload(x), ctx.alloc(4 * S * w3 * h3)
    s1, s2 = ctx.alloc(4 * S * w1 * h1 * n1), ctx.alloc(4 * S * w2 * h2 * n2)
    net.forward_fused(mi, mo, w, h, S, s1, s2)
    got = ctx.read(mo, (S, h3, w3))
    assert np.abs(e3).max() > 0.05, "degenerate test input"
    assert np.abs(got - e3).max() <= 1e-4
    mg, tgt = ctx.upload(gt), ctx.alloc(4)
    ctx.squared_error(mg, mo, tgt, w, h, w3, h3, S)
    sse_gpu = float(ctx.read(tgt, (1,))[0])
    sse_ref = port.squared_error(gt, e3, w, h, w3, h3, S)
    assert abs(psnr(sse_gpu, e3.size) - psnr(sse_ref, e3.size)) <= 0.01
    for m in (mi, mo, s1, s2, mg, tgt):
        ctx.release(m)


def _fused_vs_oracle(ctx, port, params, x, w, h, n1=64, n2=32):
    f1, f2, f3 = 9, 1, 5
    on = NetState(n1, n2, f1, f2, f3, params)
    _, _, e3 = port.net_forward(on, x, w, h, 1)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    mi, mo = ctx.upload(x), ctx.alloc(4 * w3 * h3)
    net.forward_fused(mi, mo, w, h, 1)
    got = ctx.read(mo, (1, h3, w3))
    ctx.release(mi)
    ctx.release(mo)
    return got, e3


@pytest.mark.parametrize("n1,n2", [(64, 32), (128, 64)])
def test_fused_fp16_domain_fallback(ctx, port, n1, n2):
    """The default fused kernel splits operands into FP16 halves with scales that assume
    |input| < 64.  Inputs outside that domain must still give the reference's result: the
    kernel flags them on the device and the kernel launched behind it (3xTF32 for 64/32, FP32
    SIMT for the wide network) redoes the launch.
    Also: inputs near the domain edge, and parameters far from O(1), stay within a RELATIVE
    1e-5 of the oracle (the absolute 1e-4 of north_star is for luma-range data)."""
    w, h = 200, 90
    rng = np.random.default_rng(64)
    params = make_params(rng, n1, n2, 9, 1, 5)
    x = luma_image(rng, h, w)

    def close(got, exp):
        scale = max(1.0, float(np.abs(exp).max()))
        assert np.isfinite(got).all()
        assert float(np.abs(got - exp).max()) <= 1e-5 * scale + 1e-6

    # one pixel far outside the domain, in the last strip / last rows
    x1 = x.copy()
    x1[h - 1, w - 1] = 1000.0
    close(*_fused_vs_oracle(ctx, port, params, x1, w, h, n1, n2))
    # everything outside the domain
    close(*_fused_vs_oracle(ctx, port, params, (x * 400.0 - 100.0).astype(np.float32), w, h, n1, n2))
    # inside the domain but 60x the luma range (FP16 path, large activations)
    close(*_fused_vs_oracle(ctx, port, params, (x * 120.0 - 60.0).astype(np.float32), w, h, n1, n2))
    # tiny inputs: the low halves of the split are FP16 subnormals (absolute error 2^-25 of the
    # scaled operand), the result is bias-dominated and must still match
    close(*_fused_vs_oracle(ctx, port, params, (x * 1e-4).astype(np.float32), w, h, n1, n2))
    # the reference's own initialisation scale, N(0, 0.001) (example_config.json:12-29), and
    # large parameters
    for k in (1e-3 / 0.11, 30.0):
        p2 = {name: (v * k).astype(np.float32) for name, v in params.items()}
        got, exp = _fused_vs_oracle(ctx, port, p2, x, w, h, n1, n2)
        assert np.isfinite(got).all()
        assert float(np.abs(got - exp).max()) <= 1e-5 * float(np.abs(exp).max()) + 1e-9


@pytest.mark.parametrize("n1,n2", [(64, 32), (128, 64)])
def test_fused_operand_cache_follows_the_parameters(ctx, port, n1, n2):
    """The fused kernel's packed operand image is cached per context; every way the
    parameters can change on the device must invalidate it: a host write, a device copy, a
    fill, and the training update."""
    f1, f2, f3 = 9, 1, 5
    w, h = 150, 40
    rng = np.random.default_rng(77)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x = luma_image(rng, h, w)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    mi, mo = ctx.upload(x), ctx.alloc(4 * w3 * h3)

    def check(p):
        on = NetState(n1, n2, f1, f2, f3, p)
        _, _, e3 = port.net_forward(on, x, w, h, 1)
        for _ in range(2):   # second call: served from the cache
            net.forward_fused(mi, mo, w, h, 1)
            got = ctx.read(mo, (1, h3, w3))
            assert float(np.abs(got - e3).max()) <= 1e-4

    check(params)
    # host write into one layer
    p2 = dict(params)
    p2["w2"] = (params["w2"] * 1.5).astype(np.float32)
    ctx.write(net.c.w[1], p2["w2"])
    check(p2)
    # device-to-device copy into another
    p3 = dict(p2)
    p3["b1"] = (params["b1"] + 0.05).astype(np.float32)
    tmp = ctx.upload(p3["b1"])
    ctx.copy(tmp, net.c.b[0])
    check(p3)
    # fill
    p4 = dict(p3)
    p4["b2"] = np.full_like(params["b2"], 0.02)
    ctx.fill_float(net.c.b[1], 0.02)
    check(p4)
    # one training step changes all six buffers
    px, pgt = patches(rng, 3, 33, 33)
    work = ctx.alloc(net.train_workspace_bytes(33, 33, 3))
    net.train_chunk(ctx.upload(px), ctx.upload(pgt), 33, 33, 3, work)
    net.update_all(3, 0.9, 0.001, np.array([1e-2, 1e-2, 1e-3], np.float32))
    check(net.params())


def test_inference_random_shapes(ctx, port):
    """Seeded sweep over ragged shapes of the fused 9-1-5 64/32 path: strip / band / batch
    boundaries at arbitrary positions (1x1 outputs, one-column strips, S > 1 both as virtual
    image and as separate images)."""
    n1, n2, f1, f2, f3 = 64, 32, 9, 1, 5
    rng = np.random.default_rng(20261018)
    params = make_params(rng, n1, n2, f1, f2, f3)
    on = NetState(n1, n2, f1, f2, f3, params)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    shapes = [(13, 13, 1), (137, 13, 1), (13, 137, 2), (136, 40, 1), (261, 30, 1), (260, 31, 3)]
    for _ in range(14):
        shapes.append((int(rng.integers(13, 700)), int(rng.integers(13, 160)), int(rng.integers(1, 4))))
    for (w, h, S) in shapes:
        x = np.stack([luma_image(rng, h, w) for _ in range(S)])
        _, _, e3 = port.net_forward(on, x, w, h, S)
        (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
        mi, mo = ctx.upload(x), ctx.alloc(4 * S * w3 * h3)
        net.forward_fused(mi, mo, w, h, S)
        got = ctx.read(mo, (S, h3, w3))
        assert float(np.abs(got - e3).max()) <= 1e-4, (w, h, S)
        ctx.release(mi)
        ctx.release(mo)


def test_full_size_4096_properties(ctx, port):
    """BASELINE config C3 (4096x4096, 9-1-5 64/32) at full size, through size-independent
    properties: (1) random 48x48 output windows equal the oracle run on just their receptive
    field; (2) row-band results (the multi-GPU partition, through the HOST entry point) are
    bit-identical to the single-launch result; (3) translation: cropping the input by (dy,dx)
    shifts the output by the same amount, bit for bit."""
    n1, n2, f1, f2, f3 = 64, 32, 9, 1, 5
    w = h = 4096
    rng = np.random.default_rng(4096)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x = luma_image(rng, h, w)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    if not net.fused_supported():
        pytest.skip("no fused instantiation yet")
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    pad = net.padding
    mi, mo = ctx.upload(x), ctx.alloc(4 * w3 * h3)
    net.forward_fused(mi, mo, w, h, 1)
    full = ctx.read(mo, (h3, w3))
    on = NetState(n1, n2, f1, f2, f3, params)
    for _ in range(12):
        y0, x0 = int(rng.integers(0, h3 - 48)), int(rng.integers(0, w3 - 48))
        crop = np.ascontiguousarray(x[y0:y0 + 48 + pad, x0:x0 + 48 + pad])
        _, _, e3 = port.net_forward(on, crop, 48 + pad, 48 + pad, 1)
        assert np.abs(full[y0:y0 + 48, x0:x0 + 48] - e3[0]).max() <= 1e-4
    # corners and edges too
    for (y0, x0) in ((0, 0), (h3 - 48, w3 - 48), (0, w3 - 48), (h3 - 48, 0)):
        crop = np.ascontiguousarray(x[y0:y0 + 48 + pad, x0:x0 + 48 + pad])
        _, _, e3 = port.net_forward(on, crop, 48 + pad, 48 + pad, 1)
        assert np.abs(full[y0:y0 + 48, x0:x0 + 48] - e3[0]).max() <= 1e-4
    # (2) row bands through the host entry point, N = 8 and an uneven N = 3
    for world in (8, 3):
        out = np.zeros((h3, w3), np.float32)
        for (r0, r1) in pkg.row_bands(h3, world):
            if r1 > r0:
                net.infer_rows_host(x, w, h, r0, r1, out[r0:r1])
        np.testing.assert_array_equal(out, full)
    # (3) translation
    dy, dx = 37, 101
    xs = np.ascontiguousarray(x[dy:, dx:])
    hs, ws_ = xs.shape
    m2, o2 = ctx.upload(xs), ctx.alloc(4 * (ws_ - pad) * (hs - pad))
    net.forward_fused(m2, o2, ws_, hs, 1)
    np.testing.assert_array_equal(ctx.read(o2, (hs - pad, ws_ - pad)), full[dy:, dx:])


def test_wide_network_1080p_properties(ctx, port):
    """BASELINE config C5's network (9-1-5, n1=128, n2=64) on a 1920x1080 frame through the
    wide tensor-core kernel: random 40x40 output windows equal the oracle run on their receptive
    field; row bands through the HOST entry point (sub-bands on two streams, shared operand
    image) are bit-identical to the single launch; so is a translated crop."""
    n1, n2, f1, f2, f3 = 128, 64, 9, 1, 5
    w, h = 1920, 1080
    rng = np.random.default_rng(1080)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x = luma_image(rng, h, w)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    pad = net.padding
    mi, mo = ctx.upload(x), ctx.alloc(4 * w3 * h3)
    net.forward_fused(mi, mo, w, h, 1)
    full = ctx.read(mo, (h3, w3))
    on = NetState(n1, n2, f1, f2, f3, params)
    spots = [(0, 0), (h3 - 40, w3 - 40), (0, w3 - 40), (h3 - 40, 0)]
    spots += [(int(rng.integers(0, h3 - 40)), int(rng.integers(0, w3 - 40))) for _ in range(4)]
    for (y0, x0) in spots:
        crop = np.ascontiguousarray(x[y0:y0 + 40 + pad, x0:x0 + 40 + pad])
        _, _, e3 = port.net_forward(on, crop, 40 + pad, 40 + pad, 1)
        assert np.abs(full[y0:y0 + 40, x0:x0 + 40] - e3[0]).max() <= 1e-4
    for world in (1, 3):
        out = np.zeros((h3, w3), np.float32)
        for (r0, r1) in pkg.row_bands(h3, world):
            net.infer_rows_host(x, w, h, r0, r1, out[r0:r1])
        np.testing.assert_array_equal(out, full)
    dy, dx = 11, 29
    xs = np.ascontiguousarray(x[dy:, dx:])
    hs, ws_ = xs.shape
    m2, o2 = ctx.upload(xs), ctx.alloc(4 * (ws_ - pad) * (hs - pad))
    net.forward_fused(m2, o2, ws_, hs, 1)
    np.testing.assert_array_equal(ctx.read(o2, (hs - pad, ws_ - pad)), full[dy:, dx:])
    for m in (mi, mo, m2, o2):
        ctx.release(m)


def test_host_row_pipeline_replay(ctx, port):
    """srcnn_infer_rows_host replays a captured CUDA graph while its arguments repeat: the
    replay must see new input DATA in the same host buffer, new parameter VALUES in the same
    device buffers, and a different band must not reuse the old graph."""
    n1, n2, f1, f2, f3 = 64, 32, 9, 1, 5
    w, h = 200, 1100                      # >= 1024 output rows: the multi-stream pipeline
    rng = np.random.default_rng(1100)
    params = make_params(rng, n1, n2, f1, f2, f3)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    hin, hout = pkg.PinnedBuffer((h, w)), pkg.PinnedBuffer((h3, w3))
    mi, mo = ctx.alloc(4 * w * h), ctx.alloc(4 * w3 * h3)

    def single_launch(x):
        ctx.write(mi, x)
        net.forward_fused(mi, mo, w, h, 1)
        return ctx.read(mo, (h3, w3))

    for rep in range(3):                  # capture, then two replays with new data
        x = luma_image(rng, h, w)
        hin.array[:] = x
        hout.array[:] = -1.0
        net.infer_rows_host(hin.array, w, h, 0, h3, hout.array)
        np.testing.assert_array_equal(hout.array, single_launch(x))
    # new parameter values in the same buffers
    ctx.write(net.c.w[1], (params["w2"] * 0.5).astype(np.float32))
    net.infer_rows_host(hin.array, w, h, 0, h3, hout.array)
    np.testing.assert_array_equal(hout.array, single_launch(x))
    # a different band of the same image
    out2 = np.zeros((h3 - 40, w3), np.float32)
    net.infer_rows_host(hin.array, w, h, 40, h3, out2)
    np.testing.assert_array_equal(out2, single_launch(x)[40:])
    ctx.release(mi)
    ctx.release(mo)


def test_exposed_parameters_are_never_cached(ctx, port):
    """a parameter buffer whose raw pointer left the device layer (srcnn_mem_ptr) can
    be written behind its back -- here by a device copy through a WRAPPED alias, which the
    write tracking cannot attribute to the parameter handle.  Such a buffer must not be served
    from the operand cache; srcnn_invalidate_params covers writers that never took the pointer
    through this context."""
    n1, n2, f1, f2, f3 = 64, 32, 9, 1, 5
    w, h = 150, 40
    rng = np.random.default_rng(78)
    params = make_params(rng, n1, n2, f1, f2, f3)
    x = luma_image(rng, h, w)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    mi, mo = ctx.upload(x), ctx.alloc(4 * w3 * h3)

    def run():
        net.forward_fused(mi, mo, w, h, 1)
        return ctx.read(mo, (1, h3, w3))

    def expect(p):
        _, _, e3 = port.net_forward(NetState(n1, n2, f1, f2, f3, p), x, w, h, 1)
        return e3

    assert float(np.abs(run() - expect(params)).max()) <= 1e-4
    run()                                   # cached now
    # write w2 through an alias of its raw pointer
    p2 = dict(params)
    p2["w2"] = (params["w2"] * 0.25).astype(np.float32)
    alias = ctx.wrap(ctx.mem_ptr(net.c.w[1]), 4 * p2["w2"].size)   # mem_ptr marks w2 as exposed
    src = ctx.upload(p2["w2"])
    ctx.copy(src, alias)
    assert float(np.abs(run() - expect(p2)).max()) <= 1e-4
    # and once more through the alias: still followed
    p3 = dict(p2)
    p3["w2"] = (params["w2"] * 2.0).astype(np.float32)
    ctx.write(src, p3["w2"])
    ctx.copy(src, alias)
    assert float(np.abs(run() - expect(p3)).max()) <= 1e-4
    ctx.invalidate_params()                 # harmless on top
    assert float(np.abs(run() - expect(p3)).max()) <= 1e-4


def test_gather_and_frames_host_entries(ctx, port):
    """srcnn_gather (the sample gather of execute_batch as one launch) and
    srcnn_infer_frames_host (many frames through a ring of device slots: upload / forward /
    download overlap) against their definitions."""
    rng = np.random.default_rng(91)
    # gather: 37 buffers of 33*33 floats (a size that is not a multiple of 16 bytes)
    n, per = 37, 33 * 33
    parts = [rng.normal(0, 1, per).astype(np.float32) for _ in range(n)]
    handles = [ctx.upload(p) for p in parts]
    dst = ctx.alloc(4 * per * n)
    ctx.gather(handles[::-1], 4 * per, dst)
    np.testing.assert_array_equal(ctx.read(dst, (n, per)), np.stack(parts[::-1]))
    with pytest.raises(pkg.SrcnnError):
        ctx.gather(handles, 4 * per, ctx.alloc(4 * per * (n - 1)))      # destination too small
    with pytest.raises(pkg.SrcnnError):
        ctx.gather(handles, 4 * per + 2, dst)                           # not a multiple of 4
    # frames: both tensor-core networks, more frames than one group, ragged last group
    for (n1, n2, wf, hf, nf, group) in ((128, 64, 160, 90, 7, 3), (64, 32, 96, 80, 5, 2)):
        os.environ["SRCNN_FRAMES_GROUP"] = str(group)   # read once per process: first value wins
        params = make_params(rng, n1, n2, 9, 1, 5)
        net = pkg.Net(ctx, n1, n2, 9, 1, 5, params)
        frames = np.stack([luma_image(rng, hf, wf) for _ in range(nf)])
        hin, hout = pkg.PinnedBuffer((nf, hf, wf)), pkg.PinnedBuffer((nf, hf - 12, wf - 12))
        hin.array[:] = frames
        hout.array[:] = -1
        net.infer_frames_host(hin.array, wf, hf, hout.array)
        mi, mo = ctx.upload(frames), ctx.alloc(4 * nf * (hf - 12) * (wf - 12))
        net.forward_fused(mi, mo, wf, hf, nf)
        np.testing.assert_array_equal(hout.array, ctx.read(mo, (nf, hf - 12, wf - 12)))
        _, _, e3 = port.net_forward(NetState(n1, n2, 9, 1, 5, params), frames, wf, hf, nf)
        assert float(np.abs(hout.array - e3).max()) <= 1e-4
        # a second call on the same buffers (ring state carried over correctly)
        hin.array[:] = frames[::-1]
        net.infer_frames_host(hin.array, wf, hf, hout.array)
        np.testing.assert_allclose(hout.array, e3[::-1], atol=1e-4, rtol=0)


def test_communicator_entries_on_one_rank(ctx):
    """The multi-GPU entries on a single rank: a world of 1 joins, sums are no-ops, misuse is
    reported (the 2..8-GPU behaviour is in tests/test_dp_gpu.py)."""
    c = pkg.Context(0)
    try:
        assert c.comm_info() == (0, 1)
        buf = c.upload(np.arange(8, dtype=np.float32))
        c.allreduce_sum(buf, 8)                      # no communicator: sum over one rank
        c.broadcast(buf, 8, 0)
        uid = pkg.Context.comm_unique_id()
        assert len(uid) == 128
        with pytest.raises(pkg.SrcnnError):
            c.comm_init(1, 1, uid)                   # rank outside the world
        c.comm_init(0, 1, uid)
        with pytest.raises(pkg.SrcnnError):
            c.comm_init(0, 1, uid)                   # already joined
        c.allreduce_sum(buf, 8)
        np.testing.assert_array_equal(c.read(buf, (8,)), np.arange(8, dtype=np.float32))
        with pytest.raises(pkg.SrcnnError):
            c.allreduce_sum(buf, 9)                  # outside the buffer
        c.comm_destroy()
        assert c.comm_info() == (0, 1)
    finally:
        c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,S", [((64, 32, 9, 1, 5), 150), ((64, 32, 9, 5, 5), 197),
                                   ((64, 32, 9, 5, 5), 333), ((128, 64, 9, 1, 5), 40)])
def test_no_writes_outside_the_callers_buffers(ctx, cfg, S):
    """Every buffer of a training chunk and of a fused inference call is a view into ONE arena
    with 64 KB of sentinel bytes on both sides: the tensor-core kernels (full-line staged stores,
    virtual-image tiles that hang over the last sample, the workspace carve-up of
    srcnn_train_workspace_bytes) must leave every sentinel byte alone.  (compute-sanitizer is
    closed on this pool; this is the bounds check we can run.)"""
    rng = np.random.default_rng(5)
    w = 33
    GAP = 64 * 1024
    net0 = pkg.Net(ctx, *cfg, make_params(rng, *cfg))
    sizes = {"in": 4 * S * w * w, "gt": 4 * S * w * w,   # ground truth is full size (centre crop)
             "work": net0.train_workspace_bytes(w, w, S), "grads": 4 * net0.grad_count,
             "out": 4 * S * (w - net0.padding) ** 2}
    offs, total = {}, GAP
    for k, n in sizes.items():
        offs[k] = total
        total += (n + 255) // 256 * 256 + GAP
    arena = ctx.alloc(total)
    ctx.write(arena, np.full(total, 0xA5, np.uint8))
    base = ctx.mem_ptr(arena)
    view = {k: ctx.wrap(base + offs[k], sizes[k]) for k in sizes}
    x, gt = patches(rng, S, w, w)
    ctx.write(view["in"], x)
    ctx.write(view["gt"], gt)
    ctx.write(view["grads"], np.zeros(net0.grad_count, np.float32))
    net = pkg.Net(ctx, *cfg, make_params(rng, *cfg), grad_flat=view["grads"])
    net.train_chunk(view["in"], view["gt"], w, w, S, view["work"])
    net.train_chunk(view["in"], view["gt"], w, w, S, view["work"])
    if net.fused_supported():
        net.forward_fused(view["in"], view["out"], w, w, S)
    ctx.block()
    after = ctx.read(arena, (total,), np.uint8)
    inside = np.zeros(total, bool)
    for k in sizes:
        inside[offs[k]:offs[k] + sizes[k]] = True
    bad = np.flatnonzero(~inside & (after != 0xA5))
    assert bad.size == 0, "bytes written outside the buffers: first at arena offset %d (%s)" % (
        bad[0], {k: (offs[k], sizes[k]) for k in sizes})
    assert all(np.isfinite(g).all() for g in net.grads().values())
    np.testing.assert_array_equal(ctx.read(view["in"], x.shape), x)   # inputs are read-only
    np.testing.assert_array_equal(ctx.read(view["gt"], gt.shape), gt)


@pytest.mark.gpu
def test_host_row_pipeline_async_stream_of_images(ctx, port):
    """srcnn_infer_rows_host_async: a stream of images through the two lanes (several calls queued
    before one Context.block()) gives, image by image, the bits of the blocking call; a parameter
    write between two calls waits for the calls in flight and is seen by the next one; small bands
    (single launch) and the blocking call in between work too."""
    n1, n2, f1, f2, f3 = 64, 32, 9, 1, 5
    w, h = 180, 700
    rng = np.random.default_rng(31)
    params = make_params(rng, n1, n2, f1, f2, f3)
    net = pkg.Net(ctx, n1, n2, f1, f2, f3, params)
    (_, _), (_, _), (w3, h3) = net.out_dims(w, h)
    n_img = 5
    hin = [pkg.PinnedBuffer((h, w)) for _ in range(n_img)]
    hout = [pkg.PinnedBuffer((h3, w3)) for _ in range(n_img)]
    ref = pkg.PinnedBuffer((h3, w3))
    expect = []
    for b in hin:
        b.array[:] = luma_image(rng, h, w)
        net.infer_rows_host(b.array, w, h, 0, h3, ref.array)
        expect.append(ref.array.copy())
    for rep in range(2):                  # first pass captures both lanes, second replays them
        for o in hout:
            o.array[:] = -1.0
        for i in range(n_img):
            net.infer_rows_host(hin[i].array, w, h, 0, h3, hout[i].array, block=False)
        ctx.block()
        for i in range(n_img):
            np.testing.assert_array_equal(hout[i].array, expect[i], err_msg="image %d" % i)
    # parameter write with two calls in flight, then one more call: old, old, new parameters
    net.infer_rows_host(hin[0].array, w, h, 0, h3, hout[0].array, block=False)
    net.infer_rows_host(hin[1].array, w, h, 0, h3, hout[1].array, block=False)
    ctx.write(net.c.w[1], (params["w2"] * 0.5).astype(np.float32))
    net.infer_rows_host(hin[2].array, w, h, 0, h3, hout[2].array, block=False)
    ctx.block()
    np.testing.assert_array_equal(hout[0].array, expect[0])
    np.testing.assert_array_equal(hout[1].array, expect[1])
    net.infer_rows_host(hin[2].array, w, h, 0, h3, ref.array)      # blocking, new parameters
    assert not np.array_equal(ref.array, expect[2])
    np.testing.assert_array_equal(hout[2].array, ref.array)
    # two different networks in flight: a write to the first one's parameters waits for it too
    params_b = make_params(np.random.default_rng(32), n1, n2, f1, f2, f3)
    net_b = pkg.Net(ctx, n1, n2, f1, f2, f3, params_b)
    net_b.infer_rows_host(hin[4].array, w, h, 0, h3, ref.array)
    expect_b = ref.array.copy()
    net.infer_rows_host(hin[0].array, w, h, 0, h3, ref.array)
    expect_a = ref.array.copy()
    net.infer_rows_host(hin[0].array, w, h, 0, h3, hout[0].array, block=False)
    net_b.infer_rows_host(hin[4].array, w, h, 0, h3, hout[4].array, block=False)
    ctx.write(net.c.b[0], np.zeros(n1, np.float32))        # first net, two calls ago
    ctx.block()
    np.testing.assert_array_equal(hout[0].array, expect_a)
    np.testing.assert_array_equal(hout[4].array, expect_b)
    # a band below the pipelining threshold (one launch), asynchronous, then a blocking call
    small = pkg.PinnedBuffer((100, w3))
    net.infer_rows_host(hin[3].array, w, h, 50, 150, small.array, block=False)
    net.infer_rows_host(hin[3].array, w, h, 0, h3, ref.array)      # drains the lanes first
    np.testing.assert_array_equal(small.array, ref.array[50:150])
    for b in hin + hout + [ref, small]:
        b.free()
