"""TEST INFRASTRUCTURE ONLY.

ctypes front-end for the two CPU checkers:

  * ``Oracle("port")``      -> oracle/libsrcnn_oracle.so  (our plain-C restatement,
                               oracle/srcnn_oracle.c)
  * ``Oracle("reference")`` -> oracle/_ref/libsrcnn_ref.so (the reference's own
                               OpenCL kernels compiled by g++, oracle/ref_kernels.cpp)

Both expose the same numpy-in / numpy-out methods, so a test can run the same
check against either.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product
package (cnn-super-resolution_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libsrcnn_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsrcnn_ref.so")

_f32p = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_ubyte)


def build(which="all"):
    """Build the checkers (oracle/Makefile).  'ref' needs /root/reference."""
    targets = []
    if which in ("all", "port"):
        targets.append("oracle")
    if which in ("all", "reference") and os.path.isdir("/root/reference/src/kernel"):
        targets.append("ref")
    if targets:
        subprocess.check_call(["make", "-C", HERE, "-s"] + targets)


def have(kind):
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


class _Net(C.Structure):
    _fields_ = [("n1", C.c_int), ("n2", C.c_int), ("f1", C.c_int), ("f2", C.c_int),
                ("f3", C.c_int),
                ("w", _f32p * 3), ("b", _f32p * 3),
                ("gw", _f32p * 3), ("gb", _f32p * 3),
                ("pw", _f32p * 3), ("pb", _f32p * 3)]


class _RefNet(C.Structure):
    # RefNet in ref_kernels.cpp interleaves per-kind arrays in one declaration:
    # float *w[3], *b[3], *gw[3], *gb[3], *pw[3], *pb[3];  -- same order as above.
    _fields_ = _Net._fields_


class NetState:
    """Host copy of the three layers' parameters, gradient accumulators and momentum
    state (reference: LayerAllocationPool, src/DataPipeline.hpp:11-29)."""

    def __init__(self, n1, n2, f1, f2, f3, params):
        self.n1, self.n2, self.f1, self.f2, self.f3 = n1, n2, f1, f2, f3
        shapes = self.shapes()
        self.w = [np.ascontiguousarray(params["w%d" % (i + 1)], np.float32).reshape(-1).copy()
                  for i in range(3)]
        self.b = [np.ascontiguousarray(params["b%d" % (i + 1)], np.float32).reshape(-1).copy()
                  for i in range(3)]
        for i, (k, n, f) in enumerate(shapes):
            assert self.w[i].size == f * f * k * n, (i, self.w[i].size, f * f * k * n)
            assert self.b[i].size == n
        self.gw = [np.zeros_like(x) for x in self.w]
        self.gb = [np.zeros_like(x) for x in self.b]
        self.pw = [np.zeros_like(x) for x in self.w]
        self.pb = [np.zeros_like(x) for x in self.b]

    def shapes(self):
        return [(1, self.n1, self.f1), (self.n1, self.n2, self.f2), (self.n2, 1, self.f3)]

    def cstruct(self):
        s = _Net()
        s.n1, s.n2, s.f1, s.f2, s.f3 = self.n1, self.n2, self.f1, self.f2, self.f3
        for i in range(3):
            s.w[i], s.b[i] = _fp(self.w[i]), _fp(self.b[i])
            s.gw[i], s.gb[i] = _fp(self.gw[i]), _fp(self.gb[i])
            s.pw[i], s.pb[i] = _fp(self.pw[i]), _fp(self.pb[i])
        return s

    def out_dims(self, w, h):
        w1, h1 = w - self.f1 + 1, h - self.f1 + 1
        w2, h2 = w1 - self.f2 + 1, h1 - self.f2 + 1
        w3, h3 = w2 - self.f3 + 1, h2 - self.f3 + 1
        return (w1, h1), (w2, h2), (w3, h3)


class Oracle:
    def __init__(self, kind="port"):
        assert kind in ("port", "reference")
        self.kind = kind
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.p = "oracle_" if kind == "port" else "ref_"
        L = self.lib
        self._fn("num_threads", C.c_int, [])
        self._fn("set_num_threads", None, [C.c_int])
        self._fn("forward", None, [_f32p, _f32p, _f32p, _f32p] + [C.c_int] * 7)
        self._fn("squared_error", C.c_double, [_f32p, _f32p] + [C.c_int] * 5)
        self._fn("last_layer_delta", None, [_f32p, _f32p, _f32p] + [C.c_int] * 5)
        if kind == "port":
            self._fn("deltas", None, [_f32p] * 4 + [C.c_int] * 6)
        else:
            self._fn("deltas", None, [_f32p] * 4 + [C.c_int] * 7)
        self._fn("backpropagate", None, [_f32p] * 4 + [C.c_int] * 6)
        self._fn("update_params", None, [_f32p] * 6 + [C.c_float] * 3 + [C.c_uint] * 3)
        self._fn("sum", C.c_double, [_f32p, C.c_uint, C.c_int])
        self._fn("sub_from_all", None, [_f32p, C.c_float, C.c_uint])
        netp = C.POINTER(_Net)
        self._fn("net_forward", None, [netp, _f32p] + [C.c_int] * 3 + [_f32p] * 3)
        self._fn("net_backward", None, [netp, _f32p, _f32p] + [C.c_int] * 3 + [_f32p] * 6)
        self._fn("net_update", None, [netp, C.c_uint, C.c_float, C.c_float, _f32p])
        self._fn("net_train_epoch", C.c_int if kind == "port" else None,
                 [netp, _f32p, _f32p] + [C.c_int] * 4 + [C.c_float] * 2 + [_f32p, C.c_int])
        if kind == "port":
            self._fn("subtract_mean", C.c_float, [_f32p, C.c_uint, C.c_int])
            self._fn("extract_luma", None, [_u8p, _f32p, C.c_int, C.c_int, C.c_int])
            self._fn("swap_luma", None, [_u8p, _f32p, _u8p] + [C.c_int] * 4)
            self._fn("net_validate", C.c_double, [netp, _f32p, _f32p] + [C.c_int] * 3)

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, self.p + name)
        f.restype = restype
        f.argtypes = argtypes
        setattr(self, "_" + name, f)

    # -- info -------------------------------------------------------------------
    def num_threads(self):
        return self._num_threads()

    def set_num_threads(self, n):
        self._set_num_threads(int(n))

    # -- per-kernel -------------------------------------------------------------
    def forward(self, x, W, B, k, n, f, skip_relu, in_w, in_h, S=1):
        x = np.ascontiguousarray(x, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        assert x.size == S * in_w * in_h * k and W.size >= f * f * k * n and B.size >= n
        ow, oh = in_w - f + 1, in_h - f + 1
        out = np.zeros((S, oh, ow, n), np.float32)
        self._forward(_fp(x), _fp(out), _fp(W), _fp(B), k, n, f, int(skip_relu), in_w, in_h, S)
        return out

    def squared_error(self, gt, algo, gt_w, gt_h, algo_w, algo_h, S=1):
        gt = np.ascontiguousarray(gt, np.float32)
        algo = np.ascontiguousarray(algo, np.float32)
        return float(self._squared_error(_fp(gt), _fp(algo), gt_w, gt_h, algo_w, algo_h, S))

    def last_layer_delta(self, gt, algo, gt_w, gt_h, algo_w, algo_h, S=1):
        gt = np.ascontiguousarray(gt, np.float32)
        algo = np.ascontiguousarray(algo, np.float32)
        out = np.zeros((S, algo_h, algo_w), np.float32)
        self._last_layer_delta(_fp(gt), _fp(algo), _fp(out), gt_w, gt_h, algo_w, algo_h, S)
        return out

    def deltas(self, deltas_next, layer_output, W, n_curr, f_next, n_next, out_w, out_h, S=1,
               f_curr=0):
        dn = np.ascontiguousarray(deltas_next, np.float32)
        lo = np.ascontiguousarray(layer_output, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        assert lo.size == S * out_w * out_h * n_curr
        assert dn.size == S * (out_w - f_next + 1) * (out_h - f_next + 1) * n_next
        out = np.zeros((S, out_h, out_w, n_curr), np.float32)
        if self.kind == "port":
            self._deltas(_fp(dn), _fp(lo), _fp(out), _fp(W), n_curr, f_next, n_next, out_w,
                         out_h, S)
        else:
            self._deltas(_fp(dn), _fp(lo), _fp(out), _fp(W), n_curr, f_curr, f_next, n_next,
                         out_w, out_h, S)
        return out

    def backpropagate(self, deltas, layer_input, grad_w, grad_b, n, k, f, out_w, out_h, S=1):
        """Accumulates INTO grad_w / grad_b (float32 numpy arrays), like the reference."""
        d = np.ascontiguousarray(deltas, np.float32)
        li = np.ascontiguousarray(layer_input, np.float32)
        assert d.size == S * out_w * out_h * n
        assert li.size == S * (out_w + f - 1) * (out_h + f - 1) * k
        assert grad_w.size == f * f * k * n and grad_b.size == n
        self._backpropagate(_fp(d), _fp(li), _fp(grad_w), _fp(grad_b), n, k, f, out_w, out_h, S)

    def update_params(self, w, b, gw, gb, pdw, pdb, momentum, decay, lr, batch):
        self._update_params(_fp(w), _fp(b), _fp(gw), _fp(gb), _fp(pdw), _fp(pdb),
                            momentum, decay, lr, batch, w.size, b.size)

    def sum(self, data, squared=False):
        data = np.ascontiguousarray(data, np.float32)
        return float(self._sum(_fp(data), data.size, int(squared)))

    def sub_from_all(self, data, value):
        self._sub_from_all(_fp(data), value, data.size)

    def subtract_mean(self, data, with_event=True):
        return float(self._subtract_mean(_fp(data), data.size, int(with_event)))

    def extract_luma(self, rgba, normalize=True):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        h, w = rgba.shape[:2]
        out = np.zeros((h, w), np.float32)
        self._extract_luma(rgba.ctypes.data_as(_u8p), _fp(out), w, h, int(normalize))
        return out

    def swap_luma(self, rgba, new_luma, luma_w, luma_h):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        new_luma = np.ascontiguousarray(new_luma, np.float32)
        h, w = rgba.shape[:2]
        out = np.zeros((h, w, 3), np.uint8)
        self._swap_luma(rgba.ctypes.data_as(_u8p), _fp(new_luma), out.ctypes.data_as(_u8p),
                        w, h, luma_w, luma_h)
        return out

    # -- whole net --------------------------------------------------------------
    def net_forward(self, net, x, w, h, S=1):
        x = np.ascontiguousarray(x, np.float32)
        assert x.size == S * w * h
        (w1, h1), (w2, h2), (w3, h3) = net.out_dims(w, h)
        o1 = np.zeros((S, h1, w1, net.n1), np.float32)
        o2 = np.zeros((S, h2, w2, net.n2), np.float32)
        o3 = np.zeros((S, h3, w3), np.float32)
        cs = net.cstruct()
        self._net_forward(C.byref(cs), _fp(x), w, h, S, _fp(o1), _fp(o2), _fp(o3))
        return o1, o2, o3

    def net_backward(self, net, x, gt, w, h, S, o1, o2, o3):
        x = np.ascontiguousarray(x, np.float32)
        gt = np.ascontiguousarray(gt, np.float32)
        d1, d2, d3 = np.zeros_like(o1), np.zeros_like(o2), np.zeros_like(o3)
        cs = net.cstruct()
        self._net_backward(C.byref(cs), _fp(x), _fp(gt), w, h, S, _fp(o1), _fp(o2), _fp(o3),
                           _fp(d1), _fp(d2), _fp(d3))
        return d1, d2, d3

    def net_update(self, net, batch_size, momentum, decay, lr3):
        lr = np.ascontiguousarray(lr3, np.float32)
        cs = net.cstruct()
        self._net_update(C.byref(cs), batch_size, momentum, decay, _fp(lr))

    def net_train_epoch(self, net, inputs, gts, w, h, mini_batch, momentum, decay, lr3,
                        do_update=True):
        inputs = np.ascontiguousarray(inputs, np.float32)
        gts = np.ascontiguousarray(gts, np.float32)
        n = inputs.size // (w * h)
        assert inputs.size == n * w * h and gts.size == n * w * h
        lr = np.ascontiguousarray(lr3, np.float32)
        cs = net.cstruct()
        self._net_train_epoch(C.byref(cs), _fp(inputs), _fp(gts), w, h, n, mini_batch,
                              momentum, decay, _fp(lr), int(do_update))

    def net_validate(self, net, inputs, gts, w, h):
        inputs = np.ascontiguousarray(inputs, np.float32)
        gts = np.ascontiguousarray(gts, np.float32)
        n = inputs.size // (w * h)
        cs = net.cstruct()
        return float(self._net_validate(C.byref(cs), _fp(inputs), _fp(gts), w, h, n))
