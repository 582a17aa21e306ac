"""N > 1 path on CPU: two gloo ranks shard a batch by image, exchange the 16-byte loss-statistics headers
through grouped_ssd_pytorch_b200.dist (the same helper MultiBoxLoss uses over NCCL), and their partial
losses / gradients must add up to the single-process result (the reference's DataParallel semantics:
x_max and N span the global batch).  The per-rank arithmetic is done by the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from grouped_ssd_pytorch_b200 import dist as gdist
from grouped_ssd_pytorch_b200 import synthetic as syn
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(B=6):
    pri = cases.priors("small")
    r = syn.rng(31)
    tg = syn.targets(r, B, 1, 4)
    return syn.loc(r, B, pri.shape[0]), syn.conf_logits(r, B, pri.shape[0], 2), pri, tg


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        loc, conf, pri, tg = _inputs()
        sl = gdist.shard(loc.shape[0], rank, world)
        # stage 1 on this rank's shard: local max of conf and local number of positives
        local = O.multibox_loss(loc[sl], conf[sl], pri, tg[sl], grads=False, extras=False)
        header = gdist.make_header(local["local_x_max"], local["local_n"])
        headers = gdist.all_gather_headers(header)
        assert headers.numel() == world * gdist.HEADER_BYTES
        x_max, n_total = gdist.combine_headers(headers)
        # stage 2 with the global scalars
        r = O.multibox_loss(loc[sl], conf[sl], pri, tg[sl], x_max=x_max, n_total=n_total)
        part = torch.tensor([float(r["loss_l"]), float(r["loss_c"])], dtype=torch.float64)
        dist.all_reduce(part)                      # SUM over ranks = the reference's loss
        q.put((rank, x_max, n_total, part.numpy(), r["grad_conf"], r["neg"]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_reproduce_the_global_batch_loss():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    loc, conf, pri, tg = _inputs()
    full = O.multibox_loss(loc, conf, pri, tg)
    for rank, x_max, n_total, total, grad_conf, neg in got:
        assert x_max == float(conf.max()) and n_total == int(full["num_pos"].sum())
        np.testing.assert_allclose(total, [full["loss_l"], full["loss_c"]], rtol=1e-6)
        sl = gdist.shard(loc.shape[0], rank, world)
        assert np.array_equal(neg, full["neg"][sl])                       # same hard negatives as single-process
        np.testing.assert_allclose(grad_conf, full["grad_conf"][sl], rtol=1e-6, atol=1e-12)


def test_shard_and_header_helpers():
    assert [gdist.shard(10, r, 4) for r in range(4)] == [slice(0, 3), slice(3, 6), slice(6, 9), slice(9, 10)]
    assert gdist.shard(2, 3, 4) == slice(2, 2)
    xs = np.array([-np.inf, -3.5, -0.0, 0.0, 1e-30, 2.25, np.inf], np.float32)
    o = gdist.f2ord(xs)
    assert (np.diff(o.astype(np.int64)) >= 0).all() and o[2] != o[3]      # monotone; -0 < +0
    assert np.array_equal(gdist.ord2f(o).view(np.uint32), xs.view(np.uint32))
    h = torch.cat([gdist.make_header(1.5, 7), gdist.make_header(-2.0, 5), gdist.make_header(0.25, 0)])
    assert gdist.combine_headers(h) == (1.5, 12)
    assert gdist.world()[1:] == (1, 0)


def test_peer_exchange_is_off_without_cuda_or_nccl():
    """host logic: the NVLink peer exchange is only offered for NCCL groups on CUDA devices; everything else keeps the
    all-gather of the 16-byte headers"""
    from grouped_ssd_pytorch_b200 import dist as gdist
    assert gdist.peer_exchange(None) is None          # no process group in this process
    import os
    os.environ["GSSD_PEER_XCHG"] = "0"
    try:
        assert gdist.peer_exchange(None) is None
    finally:
        del os.environ["GSSD_PEER_XCHG"]
