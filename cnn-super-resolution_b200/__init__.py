"""cnn-super-resolution_b200 -- B200 (sm_100a) SRCNN hot path behind the reference's interface.

The product is native: ``csrc/`` (hand-written CUDA kernels + the C-ABI of
``include/srcnn_b200.h``) and ``host/`` (C++ ``DataPipeline`` / ``ConfigBasedDataPipeline`` /
``LayerData`` / ``Config`` / the ``cnn`` CLI, which call only the C-ABI).  This Python module
is a thin ctypes binding over the same C-ABI, used by tests/ and bench.py and for the
torch.distributed plumbing of the multi-GPU paths.  It never computes anything itself and it
never imports oracle/: if ``libsrcnn_b200.so`` is missing or there is no CUDA device, it
raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("SRCNN_B200_LIB") or os.path.join(HERE, "libsrcnn_b200.so")
HEADER = os.path.join(ROOT, "include", "srcnn_b200.h")

NULL_MEM = 1 << 30

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

KERNEL_IDS = ["forward", "squared_err", "last_layer_delta", "deltas", "backpropagate",
              "update_params", "sum", "sub_from_all", "extract_luma", "swap_luma",
              "forward_fused", "train_fused"]


class SrcnnError(RuntimeError):
    pass


def build(verbose=False, force=False):
    """Compile csrc/ for sm_100a into libsrcnn_b200.so (in-tree, so it travels to the GPU box)."""
    srcs = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))]
    deps = srcs + [HEADER]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH, os.path.join(HERE, "csrc", "srcnn.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None

_u64 = C.c_uint64
_vp = C.c_void_p
_i = C.c_int
_f = C.c_float
_u = C.c_uint
_sz = C.c_size_t


class CNet(C.Structure):
    """struct srcnn_net of include/srcnn_b200.h"""
    _fields_ = [("n1", _i), ("n2", _i), ("f1", _i), ("f2", _i), ("f3", _i),
                ("w", _u64 * 3), ("b", _u64 * 3),
                ("grad_w", _u64 * 3), ("grad_b", _u64 * 3),
                ("prev_dw", _u64 * 3), ("prev_db", _u64 * 3)]


_SIGS = {
    "srcnn_ctx_create": (_i, [_i, _i, C.POINTER(_vp)]),
    "srcnn_ctx_create_on_stream": (_i, [_i, _vp, _i, C.POINTER(_vp)]),
    "srcnn_ctx_destroy": (_i, [_vp]),
    "srcnn_last_error": (C.c_char_p, []),
    "srcnn_block": (_i, [_vp]),
    "srcnn_device_info": (_i, [_vp, C.c_char_p, _sz, C.POINTER(_i), C.POINTER(_sz)]),
    "srcnn_profile_get": (_i, [_vp, _i, C.POINTER(_u64), C.POINTER(_u64)]),
    "srcnn_launch_count": (_i, [_vp, C.POINTER(_u64)]),
    "srcnn_stream": (_i, [_vp, C.POINTER(_vp)]),
    "srcnn_alloc": (_i, [_vp, _sz, C.POINTER(_u64)]),
    "srcnn_wrap": (_i, [_vp, _vp, _sz, C.POINTER(_u64)]),
    "srcnn_release": (_i, [_vp, _u64]),
    "srcnn_mem_size": (_i, [_vp, _u64, C.POINTER(_sz)]),
    "srcnn_mem_ptr": (_i, [_vp, _u64, C.POINTER(_vp)]),
    "srcnn_mem_usage": (_i, [_vp, C.POINTER(_sz)]),
    "srcnn_write": (_i, [_vp, _u64, _sz, _sz, _vp, _i]),
    "srcnn_read": (_i, [_vp, _u64, _sz, _sz, _vp, _i]),
    "srcnn_copy": (_i, [_vp, _u64, _u64, _sz]),
    "srcnn_copy_region": (_i, [_vp, _u64, _sz, _u64, _sz, _sz]),
    "srcnn_gather": (_i, [_vp, C.POINTER(_u64), _i, _sz, _u64]),
    "srcnn_fill_float": (_i, [_vp, _u64, _f]),
    "srcnn_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "srcnn_host_free": (_i, [_vp]),
    "srcnn_forward_layer": (_i, [_vp, _u64, _u64, _u64, _u64] + [_i] * 7),
    "srcnn_squared_error": (_i, [_vp, _u64, _u64, _u64] + [_i] * 5),
    "srcnn_last_layer_delta": (_i, [_vp, _u64, _u64, _u64] + [_i] * 5),
    "srcnn_deltas": (_i, [_vp, _u64, _u64, _u64, _u64] + [_i] * 6),
    "srcnn_backpropagate": (_i, [_vp, _u64, _u64, _u64, _u64] + [_i] * 6),
    "srcnn_update_params": (_i, [_vp] + [_u64] * 6 + [_f] * 3 + [_u] * 3),
    "srcnn_sum": (_i, [_vp, _u64, _u, _i, _u64]),
    "srcnn_sub_from_all": (_i, [_vp, _u64, _f, _u]),
    "srcnn_extract_luma": (_i, [_vp, _u64, _u64, _i, _i, _i]),
    "srcnn_swap_luma": (_i, [_vp, _u64, _u64, _u64] + [_i] * 4),
    "srcnn_forward_fused_supported": (_i, [C.POINTER(CNet)]),
    "srcnn_forward_fused": (_i, [_vp, C.POINTER(CNet), _u64, _u64, _i, _i, _i, _u64, _u64]),
    "srcnn_infer_rows_host": (_i, [_vp, C.POINTER(CNet), _vp, _i, _i, _i, _i, _vp]),
    "srcnn_infer_rows_host_async": (_i, [_vp, C.POINTER(CNet), _vp, _i, _i, _i, _i, _vp]),
    "srcnn_infer_frames_host": (_i, [_vp, C.POINTER(CNet), _vp, _i, _i, _i, _vp]),
    "srcnn_train_chunk": (_i, [_vp, C.POINTER(CNet), _u64, _u64, _i, _i, _i, _u64]),
    "srcnn_train_chunks_host": (_i, [_vp, C.POINTER(CNet), _vp, _vp, _i, _i, _i, _i, _u64]),
    "srcnn_train_chunk_buffers": (_i, [_vp, C.POINTER(CNet), _u64, _u64, _i, _i, _i,
                                       _u64, _u64, _u64, _u64, _u64, _u64]),
    "srcnn_train_workspace_bytes": (_sz, [C.POINTER(CNet), _i, _i, _i]),
    "srcnn_train_materializes_d1": (_i, [_vp, C.POINTER(CNet)]),
    "srcnn_update_all": (_i, [_vp, C.POINTER(CNet), _u, _f, _f, C.POINTER(_f)]),
    "srcnn_validate_chunk": (_i, [_vp, C.POINTER(CNet), _u64, _u64, _i, _i, _i, _u64, _u64]),
    "srcnn_invalidate_params": (_i, [_vp]),
    "srcnn_comm_unique_id": (_i, [_vp]),
    "srcnn_comm_init": (_i, [_vp, _i, _i, _vp]),
    "srcnn_comm_destroy": (_i, [_vp]),
    "srcnn_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "srcnn_allreduce_sum": (_i, [_vp, _u64, _sz, _sz]),
    "srcnn_allreduce_grads": (_i, [_vp, C.POINTER(CNet)]),
    "srcnn_broadcast": (_i, [_vp, _u64, _sz, _i]),
}


def exported_symbols():
    """Every entry point declared in include/srcnn_b200.h (checked by the CPU test-suite)."""
    return sorted(_SIGS)


def lib():
    """Load libsrcnn_b200.so.  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SrcnnError("%s is missing: run `python -c 'import __graft_entry__ as g; "
                             "g.build()'` (nvcc, sm_100a) first; there is no CPU fallback"
                             % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise SrcnnError("srcnn error %d: %s" % (rc, lib().srcnn_last_error().decode()))


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class PinnedBuffer:
    """Page-locked host staging buffer (srcnn_host_alloc) exposed as a numpy array."""

    def __init__(self, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = _vp()
        _check(lib().srcnn_host_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().srcnn_host_free(_vp(self.ptr))
            self.ptr = None


class Context:
    """One device + one in-order stream + a handle table.
    Mirrors opencl::Context (reference: src/opencl/Context.hpp:72-299)."""

    def __init__(self, device=0, profile=False, stream=None):
        L = lib()
        h = _vp()
        if stream is None:
            _check(L.srcnn_ctx_create(device, int(profile), C.byref(h)))
        else:
            _check(L.srcnn_ctx_create_on_stream(device, _vp(stream), int(profile), C.byref(h)))
        self.h = h
        self.L = L
        self.device = device

    # -- lifetime / info --
    def close(self):
        if self.h:
            self.L.srcnn_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def block(self):
        _check(self.L.srcnn_block(self.h))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, mem = _i(), _sz()
        _check(self.L.srcnn_device_info(self.h, name, 256, C.byref(sm), C.byref(mem)))
        return name.value.decode(), sm.value, mem.value

    def launch_count(self):
        n = _u64()
        _check(self.L.srcnn_launch_count(self.h, C.byref(n)))
        return n.value

    def profile(self):
        out = {}
        for i, name in enumerate(KERNEL_IDS):
            ns, n = _u64(), _u64()
            _check(self.L.srcnn_profile_get(self.h, i, C.byref(ns), C.byref(n)))
            out[name] = (ns.value, n.value)
        return out

    def stream(self):
        p = _vp()
        _check(self.L.srcnn_stream(self.h, C.byref(p)))
        return p.value

    # -- multi-GPU (NCCL communicator owned by the context) --
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL id created on rank 0; hand it to the other ranks."""
        buf = (C.c_ubyte * 128)()
        _check(lib().srcnn_comm_unique_id(C.cast(buf, _vp)))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        _check(self.L.srcnn_comm_init(self.h, rank, world, C.cast(buf, _vp)))

    def comm_destroy(self):
        _check(self.L.srcnn_comm_destroy(self.h))

    def comm_info(self):
        r, w = _i(), _i()
        _check(self.L.srcnn_comm_info(self.h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def allreduce_sum(self, mem, count, offset=0):
        _check(self.L.srcnn_allreduce_sum(self.h, mem, offset, count))

    def broadcast(self, mem, count, root=0):
        _check(self.L.srcnn_broadcast(self.h, mem, count, root))

    def invalidate_params(self):
        _check(self.L.srcnn_invalidate_params(self.h))

    # -- memory --
    def alloc(self, nbytes):
        m = _u64()
        _check(self.L.srcnn_alloc(self.h, int(nbytes), C.byref(m)))
        return m.value

    def wrap(self, device_ptr, nbytes):
        m = _u64()
        _check(self.L.srcnn_wrap(self.h, _vp(device_ptr), int(nbytes), C.byref(m)))
        return m.value

    def release(self, mem):
        _check(self.L.srcnn_release(self.h, mem))

    def mem_size(self, mem):
        s = _sz()
        _check(self.L.srcnn_mem_size(self.h, mem, C.byref(s)))
        return s.value

    def mem_ptr(self, mem):
        p = _vp()
        _check(self.L.srcnn_mem_ptr(self.h, mem, C.byref(p)))
        return p.value

    def mem_usage(self):
        s = _sz()
        _check(self.L.srcnn_mem_usage(self.h, C.byref(s)))
        return s.value

    def write(self, mem, array, offset=0, block=True):
        a = np.ascontiguousarray(array)
        _check(self.L.srcnn_write(self.h, mem, offset, a.nbytes, _np_ptr(a), int(block)))

    def read(self, mem, shape, dtype=np.float32, offset=0):
        out = np.empty(shape, dtype)
        _check(self.L.srcnn_read(self.h, mem, offset, out.nbytes, _np_ptr(out), 1))
        return out

    def upload(self, array, dtype=np.float32):
        a = np.ascontiguousarray(array, dtype)
        m = self.alloc(max(a.nbytes, 4))
        if a.nbytes:
            self.write(m, a)
        return m

    def zeros(self, nfloats):
        m = self.alloc(4 * max(int(nfloats), 1))
        self.fill_float(m, 0.0)
        return m

    def copy(self, src, dst, dst_offset=0):
        _check(self.L.srcnn_copy(self.h, src, dst, dst_offset))

    def copy_region(self, src, src_offset, dst, dst_offset, nbytes):
        _check(self.L.srcnn_copy_region(self.h, src, src_offset, dst, dst_offset, nbytes))

    def gather(self, handles, nbytes_each, dst):
        """copies the equally-sized buffers `handles` into consecutive slots of dst (one launch)"""
        arr = (_u64 * len(handles))(*handles)
        _check(self.L.srcnn_gather(self.h, arr, len(handles), int(nbytes_each), dst))

    def fill_float(self, mem, value):
        _check(self.L.srcnn_fill_float(self.h, mem, float(value)))

    # -- kernels (one per reference .cl entry point) --
    def forward_layer(self, inp, out, W, B, k, n, f, skip_relu, in_w, in_h, S=1):
        _check(self.L.srcnn_forward_layer(self.h, inp, out, W, B, k, n, f, int(skip_relu),
                                          in_w, in_h, S))

    def squared_error(self, gt, algo, target, gt_w, gt_h, algo_w, algo_h, S=1):
        _check(self.L.srcnn_squared_error(self.h, gt, algo, target, gt_w, gt_h, algo_w, algo_h, S))

    def last_layer_delta(self, gt, algo, target, gt_w, gt_h, algo_w, algo_h, S=1):
        _check(self.L.srcnn_last_layer_delta(self.h, gt, algo, target, gt_w, gt_h, algo_w,
                                             algo_h, S))

    def deltas(self, deltas_next, layer_output, target, W, n_curr, f_next, n_next, out_w, out_h,
               S=1):
        _check(self.L.srcnn_deltas(self.h, deltas_next, layer_output, target, W, n_curr, f_next,
                                   n_next, out_w, out_h, S))

    def backpropagate(self, deltas, layer_input, grad_w, grad_b, n, k, f, out_w, out_h, S=1):
        _check(self.L.srcnn_backpropagate(self.h, deltas, layer_input, grad_w, grad_b, n, k, f,
                                          out_w, out_h, S))

    def update_params(self, w, b, gw, gb, pdw, pdb, momentum, decay, lr, batch, wsize, bsize):
        _check(self.L.srcnn_update_params(self.h, w, b, gw, gb, pdw, pdb, momentum, decay, lr,
                                          batch, wsize, bsize))

    def sum(self, data, length, squared, target):
        _check(self.L.srcnn_sum(self.h, data, length, int(squared), target))

    def sub_from_all(self, data, value, length):
        _check(self.L.srcnn_sub_from_all(self.h, data, value, length))

    def extract_luma(self, rgba, target, w, h, normalize=True):
        _check(self.L.srcnn_extract_luma(self.h, rgba, target, w, h, int(normalize)))

    def swap_luma(self, rgba, new_luma, target, gt_w, gt_h, luma_w, luma_h):
        _check(self.L.srcnn_swap_luma(self.h, rgba, new_luma, target, gt_w, gt_h, luma_w, luma_h))


class Net:
    """Device-resident parameters, gradient accumulators and momentum state of the three
    layers (reference: GpuAllocationPool / LayerAllocationPool, src/DataPipeline.hpp:11-29,
    src/ConfigBasedDataPipeline.hpp:33-39) plus the fused entry points."""

    def __init__(self, ctx, n1, n2, f1, f2, f3, params, grad_flat=None):
        self.ctx = ctx
        self.n1, self.n2, self.f1, self.f2, self.f3 = n1, n2, f1, f2, f3
        c = CNet()
        c.n1, c.n2, c.f1, c.f2, c.f3 = n1, n2, f1, f2, f3
        self.sizes = []
        for l, (k, n, f) in enumerate(self.shapes()):
            ws, bs = f * f * k * n, n
            w = np.ascontiguousarray(params["w%d" % (l + 1)], np.float32).reshape(-1)
            b = np.ascontiguousarray(params["b%d" % (l + 1)], np.float32).reshape(-1)
            assert w.size == ws and b.size == bs, (l, w.size, ws, b.size, bs)
            c.w[l], c.b[l] = ctx.upload(w), ctx.upload(b)
            c.prev_dw[l], c.prev_db[l] = ctx.zeros(ws), ctx.zeros(bs)
            self.sizes.append((ws, bs))
        # the six gradient tensors live in ONE contiguous buffer so that data-parallel
        # training needs a single all-reduce; grad_flat may be caller memory
        # (a torch tensor) registered with Context.wrap.
        self.grad_count = sum(ws + bs for ws, bs in self.sizes)
        if grad_flat is None:
            self.grad_flat = ctx.zeros(self.grad_count)
        else:
            self.grad_flat = grad_flat
        base = ctx.mem_ptr(self.grad_flat)
        off = 0
        for l, (ws, bs) in enumerate(self.sizes):
            c.grad_w[l] = ctx.wrap(base + 4 * off, 4 * ws)
            off += ws
            c.grad_b[l] = ctx.wrap(base + 4 * off, 4 * bs)
            off += bs
        self.c = c

    def shapes(self):
        return [(1, self.n1, self.f1), (self.n1, self.n2, self.f2), (self.n2, 1, self.f3)]

    def out_dims(self, w, h):
        w1, h1 = w - self.f1 + 1, h - self.f1 + 1
        w2, h2 = w1 - self.f2 + 1, h1 - self.f2 + 1
        return (w1, h1), (w2, h2), (w2 - self.f3 + 1, h2 - self.f3 + 1)

    @property
    def padding(self):
        """Config::total_padding (reference: src/Config.cpp:44)"""
        return self.f1 + self.f2 + self.f3 - 3

    def fused_supported(self):
        return bool(self.ctx.L.srcnn_forward_fused_supported(C.byref(self.c)))

    def forward_fused(self, inp, out, w, h, S=1, scratch1=NULL_MEM, scratch2=NULL_MEM):
        _check(self.ctx.L.srcnn_forward_fused(self.ctx.h, C.byref(self.c), inp, out, w, h, S,
                                              scratch1, scratch2))

    def infer_rows_host(self, host_in, w, h, row0, row1, host_out, block=True):
        """host_in: full [h][w] float32 image; host_out: array whose first row is output row row0.
        block=False queues the call (two can be in flight); Context.block() completes them."""
        assert host_in.dtype == np.float32 and host_out.dtype == np.float32
        fn = self.ctx.L.srcnn_infer_rows_host if block else self.ctx.L.srcnn_infer_rows_host_async
        _check(fn(self.ctx.h, C.byref(self.c), _np_ptr(host_in), w, h, row0, row1,
                  _np_ptr(host_out)))

    def infer_frames_host(self, host_in, w, h, host_out):
        """host_in: [n][h][w] float32 frames; host_out: [n][h3][w3].  Upload / forward / download
        of consecutive frame groups overlap."""
        assert host_in.dtype == np.float32 and host_out.dtype == np.float32
        assert host_in.flags.c_contiguous and host_out.flags.c_contiguous
        _check(self.ctx.L.srcnn_infer_frames_host(self.ctx.h, C.byref(self.c), _np_ptr(host_in), w,
                                                  h, int(host_in.shape[0]), _np_ptr(host_out)))

    def train_workspace_bytes(self, w, h, S):
        return self.ctx.L.srcnn_train_workspace_bytes(C.byref(self.c), w, h, S)

    def train_materializes_d1(self):
        """False when the chunk entries keep the layer-1 deltas inside the gradient kernel."""
        return bool(self.ctx.L.srcnn_train_materializes_d1(self.ctx.h, C.byref(self.c)))

    def train_chunk(self, inp, gt, w, h, S, work):
        _check(self.ctx.L.srcnn_train_chunk(self.ctx.h, C.byref(self.c), inp, gt, w, h, S, work))

    def train_chunks_host(self, host_in, host_gt, w, h, chunk, work):
        """Forward + backward over HOST sample arrays [n][h][w] (pinned for overlap), uploads
        pipelined with the training of the previous chunk; gradients accumulate."""
        assert host_in.dtype == np.float32 and host_gt.dtype == np.float32
        assert host_in.shape == host_gt.shape and host_in.flags.c_contiguous
        _check(self.ctx.L.srcnn_train_chunks_host(self.ctx.h, C.byref(self.c), _np_ptr(host_in),
                                                  _np_ptr(host_gt), w, h, int(host_in.shape[0]),
                                                  chunk, work))

    def update_all(self, batch_size, momentum, decay, lr3):
        lr = (C.c_float * 3)(*[float(v) for v in lr3])
        _check(self.ctx.L.srcnn_update_all(self.ctx.h, C.byref(self.c), batch_size, momentum,
                                           decay, lr))

    def validate_chunk(self, inp, gt, w, h, S, work, target):
        _check(self.ctx.L.srcnn_validate_chunk(self.ctx.h, C.byref(self.c), inp, gt, w, h, S,
                                               work, target))

    def allreduce_grads(self):
        """Sum the six gradient accumulators over the ranks of the context's communicator."""
        _check(self.ctx.L.srcnn_allreduce_grads(self.ctx.h, C.byref(self.c)))

    def params(self):
        out = {}
        for l, (ws, bs) in enumerate(self.sizes):
            out["w%d" % (l + 1)] = self.ctx.read(self.c.w[l], (ws,))
            out["b%d" % (l + 1)] = self.ctx.read(self.c.b[l], (bs,))
        return out

    def grads(self):
        flat = self.ctx.read(self.grad_flat, (self.grad_count,))
        out, off = {}, 0
        for l, (ws, bs) in enumerate(self.sizes):
            out["w%d" % (l + 1)] = flat[off:off + ws]
            off += ws
            out["b%d" % (l + 1)] = flat[off:off + bs]
            off += bs
        return out


# ----------------------------------------------------------------------------- sharding
def row_bands(out_h, world_size):
    """Row-band partition of the OUTPUT rows of one image across ranks: rank g gets
    [g*ceil(out_h/N), ...).  Each band needs `halo` = f1+f2+f3-3 extra INPUT rows, no exchange."""
    per = -(-out_h // world_size)
    bands = []
    for g in range(world_size):
        r0 = min(g * per, out_h)
        r1 = min(r0 + per, out_h)
        bands.append((r0, r1))
    return bands


def patch_shards(n_patches, world_size):
    """Contiguous even split of the training patches across ranks (data parallel)."""
    base, rem = divmod(n_patches, world_size)
    out, start = [], 0
    for g in range(world_size):
        cnt = base + (1 if g < rem else 0)
        out.append((start, start + cnt))
        start += cnt
    return out
