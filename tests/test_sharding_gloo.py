"""N > 1 host-side logic on CPU: world_size-2 `gloo` process groups run the SAME sharding /
exchange code paths bench.py uses on NCCL (pkg.row_bands, pkg.patch_shards, one all-reduce of
the flat gradient, update with the GLOBAL batch size), with the oracle standing in for the
device kernels.  Checks:
  * training: sum over ranks of per-shard gradients == single-process gradient, and the
    parameters after the update are identical on every rank and equal to the 1-process run;
  * inference: row bands with an (f1+f2+f3-3)-row halo, no exchange, reassemble bit-exactly.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _pkg
from helpers import luma_image, make_params, patches
from oracle.loader import NetState, Oracle

pkg = _pkg.load()
CFG = (8, 4, 9, 1, 5)
LR = np.array([1e-3, 1e-3, 1e-4], np.float32)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def flat_grad(net):
    return np.concatenate([np.concatenate([net.gw[l], net.gb[l]]) for l in range(3)])


def set_flat_grad(net, flat):
    off = 0
    for l in range(3):
        for arr in (net.gw[l], net.gb[l]):
            arr[:] = flat[off:off + arr.size]
            off += arr.size


def worker(rank, world, port_no, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle("port")
    orc.set_num_threads(2)
    rng = np.random.default_rng(11)
    params = make_params(rng, *CFG)
    n, w = 10, 33
    x, gt = patches(rng, n, w, w)
    # ---- training: data-parallel shard + ONE all-reduce + update with the global batch
    net = NetState(*CFG, params)
    a, b = pkg.patch_shards(n, world)[rank]
    orc.net_train_epoch(net, x[a:b], gt[a:b], w, w, 4, 0.9, 0.001, LR, do_update=False)
    g = torch.from_numpy(flat_grad(net))
    dist.all_reduce(g)
    set_flat_grad(net, g.numpy())
    orc.net_update(net, n, 0.9, 0.001, LR)
    # ---- inference: my row band only
    img = luma_image(np.random.default_rng(12), 61, 47)
    h3 = 61 - 12
    r0, r1 = pkg.row_bands(h3, world)[rank]
    band = np.ascontiguousarray(img[r0:r1 + 12])
    inf = NetState(*CFG, params)
    _, _, o3 = orc.net_forward(inf, band, 47, band.shape[0], 1)
    np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), grad=g.numpy(),
             w=np.concatenate(net.w), b=np.concatenate(net.b), band=o3[0], rows=np.array([r0, r1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_data_parallel_and_row_bands_match_single_process(tmp_path, world):
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    orc = Oracle("port")
    rng = np.random.default_rng(11)
    params = make_params(rng, *CFG)
    n, w = 10, 33
    x, gt = patches(rng, n, w, w)
    ref = NetState(*CFG, params)
    orc.net_train_epoch(ref, x, gt, w, w, 4, 0.9, 0.001, LR, do_update=False)
    g1 = flat_grad(ref)
    orc.net_update(ref, n, 0.9, 0.001, LR)
    img = luma_image(np.random.default_rng(12), 61, 47)
    full = orc.net_forward(NetState(*CFG, params), img, 47, 61, 1)[2][0]
    outs = [dict(np.load(tmp_path / ("rank%d.npz" % r))) for r in range(world)]
    assembled = np.zeros_like(full)
    for o in outs:
        np.testing.assert_allclose(o["grad"], g1, rtol=2e-5, atol=1e-6)
        np.testing.assert_array_equal(o["w"], outs[0]["w"])          # replicas stay identical
        np.testing.assert_allclose(o["w"], np.concatenate(ref.w), rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(o["b"], np.concatenate(ref.b), rtol=1e-5, atol=1e-8)
        r0, r1 = o["rows"]
        assembled[r0:r1] = o["band"]
    np.testing.assert_array_equal(assembled, full)
