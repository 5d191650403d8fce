"""Data-parallel correctness on real GPUs (needs >= 2 devices: `gpurun --gpus 2`; skipped on a
one-GPU box).  One process per GPU, each with its own context; the communicator is the one the
C-ABI owns (srcnn_comm_init -> NCCL over NVLink), no torch anywhere.

What must hold:
  * N ranks x srcnn_allreduce_grads x srcnn_update_all(GLOBAL batch) == the 1-GPU parameters to
    summation-order tolerance, and == the reference's own kernels (committed fixture) to 1e-4;
  * the validation squared error summed with the 1-float all-reduce == the 1-GPU value
    (reference: src/ConfigBasedDataPipeline.cpp:177-187, src/Main_cl.cpp:174-192).
"""
import ctypes as C
import multiprocessing as mp
import os

import numpy as np
import pytest

import _pkg
from conftest import load_npz

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        rt = C.CDLL("libcudart.so")
    except OSError:
        try:
            rt = C.CDLL("libcudart.so.12")
        except OSError:
            return 0
    n = C.c_int(0)
    return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0


def _inputs(g):
    from helpers import make_params, patches
    cfg = tuple(int(v) for v in g["cfg"])
    ns, w, h, chunk, epochs = (int(v) for v in g["dims"])
    rng = np.random.default_rng(int(g["seed"]))
    params = make_params(rng, *cfg)
    x, gt = patches(rng, ns, w, h)
    return cfg, (ns, w, h, chunk, epochs), params, x, gt


def _worker(rank, world, name, uid_q, out_q):
    """One rank: its shard of the patches, the product's own communicator."""
    try:
        pkg = _pkg.load()
        g = load_npz(name)
        cfg, (ns, w, h, chunk, epochs), params, x, gt = _inputs(g)
        ctx = pkg.Context(rank)
        if rank == 0:
            uid = pkg.Context.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        ctx.comm_init(rank, world, uid)
        assert ctx.comm_info() == (rank, world)
        net = pkg.Net(ctx, *cfg, params)
        # every rank starts from rank 0's parameters (here they are equal already; the call is
        # what a launcher does after a random initialisation)
        for l in range(3):
            ctx.broadcast(net.c.w[l], net.sizes[l][0], 0)
            ctx.broadcast(net.c.b[l], net.sizes[l][1], 0)
        s0, s1 = pkg.patch_shards(ns, world)[rank]
        lx, lgt = np.ascontiguousarray(x[s0:s1]), np.ascontiguousarray(gt[s0:s1])
        lchunk = min(chunk, s1 - s0)
        work = ctx.alloc(net.train_workspace_bytes(w, h, lchunk))
        res = {}
        for e in range(1, epochs + 1):
            net.train_chunks_host(lx, lgt, w, h, lchunk, work)
            net.allreduce_grads()                      # the ONE exchange step of the path
            net.update_all(ns, float(g["momentum"]), float(g["decay"]), g["lr"])
            res["e%d" % e] = net.params()
        # validation pass over the shard + 1-float all-reduce
        tgt = ctx.alloc(4)
        mi, mg = ctx.upload(lx), ctx.upload(lgt)
        vwork = ctx.alloc(net.train_workspace_bytes(w, h, s1 - s0))
        net.validate_chunk(mi, mg, w, h, s1 - s0, vwork, tgt)
        res["sse_local"] = float(ctx.read(tgt, (1,))[0])
        ctx.allreduce_sum(tgt, 1)
        res["sse"] = float(ctx.read(tgt, (1,))[0])
        ctx.comm_destroy()
        ctx.close()
        out_q.put((rank, res))
    except Exception as exc:   # surface the failure in the parent
        import traceback
        out_q.put((rank, "ERROR: %s\n%s" % (exc, traceback.format_exc())))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_data_parallel_equals_single_gpu(world):
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    name = "ref_train_c2_epoch.npz"
    pkg = _pkg.load()
    g = load_npz(name)
    cfg, (ns, w, h, chunk, epochs), params, x, gt = _inputs(g)

    # spawned ranks re-import this module: they need the repo root and tests/ on their path
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join(
        [root, os.path.join(root, "tests")] + [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p])
    mpc = mp.get_context("spawn")
    uid_q, out_q = mpc.Queue(), mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, name, uid_q, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, res = out_q.get(timeout=600)
        assert not isinstance(res, str), "rank %d: %s" % (rank, res)
        results[rank] = res
    for p in procs:
        p.join(timeout=60)

    # the single-GPU run of the same epochs
    with pkg.Context(0) as ctx:
        net = pkg.Net(ctx, *cfg, params)
        work = ctx.alloc(net.train_workspace_bytes(w, h, chunk))
        single = {}
        for e in range(1, epochs + 1):
            net.train_chunks_host(x, gt, w, h, chunk, work)
            net.update_all(ns, float(g["momentum"]), float(g["decay"]), g["lr"])
            single["e%d" % e] = net.params()
        tgt, mi, mg = ctx.alloc(4), ctx.upload(x), ctx.upload(gt)
        vwork = ctx.alloc(net.train_workspace_bytes(w, h, ns))
        net.validate_chunk(mi, mg, w, h, ns, vwork, tgt)
        single_sse = float(ctx.read(tgt, (1,))[0])

    for e in range(1, epochs + 1):
        for key in single["e%d" % e]:
            ref = g["e%d_%s" % (e, key)]
            one = single["e%d" % e][key]
            for r in range(world):
                got = results[r]["e%d" % e][key]
                # replicas stay bit-identical: same all-reduced gradient, same update
                np.testing.assert_array_equal(got, results[0]["e%d" % e][key],
                                              err_msg="rank %d diverged from rank 0 (%s)" % (r, key))
            got = results[0]["e%d" % e][key]
            # DP == 1 GPU up to the summation order of the gradient
            np.testing.assert_allclose(got, one, rtol=2e-5, atol=1e-8, err_msg="DP vs 1 GPU " + key)
            # DP == the reference's own kernels (north_star: 1e-4 relative per epoch)
            np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-7, err_msg="DP vs reference " + key)
    total = sum(results[r]["sse_local"] for r in range(world))
    for r in range(world):
        assert results[r]["sse"] == pytest.approx(total, rel=1e-6)
        assert results[r]["sse"] == pytest.approx(single_sse, rel=1e-5)
