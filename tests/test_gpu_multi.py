"""Two GPUs, one process each (NCCL): batch-sharded MultiBoxLoss with the 16-byte statistics all-gather
must reproduce the single-GPU loss, masks and gradients; Detect shards trivially."""
import os
import socket

import numpy as np
import pytest
import torch

import cases
from grouped_ssd_pytorch_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(B=8):
    B = int(os.environ.get("GSSD_TEST_MULTI_B", B))
    pri = cases.priors("v2")
    r = syn.rng(77)
    tg = syn.targets(r, B, 1, 5)
    return syn.loc(r, B, pri.shape[0]), syn.conf_logits(r, B, pri.shape[0], 2), pri, tg


def _run(loc, conf, pri, tg, dev, steps=1):
    """`steps` calls of the criterion on the same inputs (the exchange's epoch advances every call); returns the last"""
    from grouped_ssd_pytorch_b200.layers import MultiBoxLoss
    crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
    crit.keep_masks = True
    for k in range(steps + (1 if steps > 1 else 0)):
        if k == steps:
            # one more call through a SECOND criterion object on the same stream: its exchange starts at epoch 1 while the
            # launch counter of the shared local state is at `steps` (bench.py's N-rank check found the two mixed up)
            crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
            crit.keep_masks = True
        l = torch.from_numpy(loc).to(dev).requires_grad_()
        c = torch.from_numpy(conf).to(dev).requires_grad_()
        ll, lc = crit((l, c, torch.from_numpy(pri).to(dev)), [torch.from_numpy(t).to(dev) for t in tg])
        (ll + lc).backward()
    return ll, lc, l.grad, c.grad, crit.last_masks


def _worker(rank, world, port, q, xchg):
    os.environ["GSSD_PEER_XCHG"] = xchg
    import torch.distributed as dist
    from grouped_ssd_pytorch_b200 import dist as gdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        loc, conf, pri, tg = _inputs()
        sl = gdist.shard(loc.shape[0], rank, world)
        ll, lc, gl, gc, m = _run(loc[sl], conf[sl], pri, tg[sl], dev, steps=3)
        tot = torch.stack([ll.detach(), lc.detach()]).double()
        dist.all_reduce(tot)
        # the native host-buffer pipeline in its two-phase (begin / all-gather / finish) form gives the same step
        from grouped_ssd_pytorch_b200.pipeline import HostPipeline
        pipe = HostPipeline(sl.stop - sl.start, torch.from_numpy(pri).to(dev), depth=2, device=dev)
        hb = pipe.host_buffers()
        hb.loc.copy_(torch.from_numpy(loc[sl])); hb.conf.copy_(torch.from_numpy(conf[sl]))
        for _ in range(3):
            t = pipe.submit(hb, [torch.from_numpy(x) for x in tg[sl]], detect=False)
            pipe.wait(t)
        pgl, pgc = pipe.grads(t)
        pipe_ok = bool(float(hb.losses[0]) == float(ll.detach()) and float(hb.losses[1]) == float(lc.detach())
                       and torch.equal(pgl, gl) and torch.equal(pgc, gc))
        pipe.close()
        q.put((rank, tot.cpu().numpy(), gl.cpu().numpy(), gc.cpu().numpy(), m["neg"].cpu().numpy(), pipe_ok))
        used = gdist.peer_exchange(None) is not None
        assert used == (xchg == "1"), "peer exchange in use: %s, requested: %s" % (used, xchg)
        dist.barrier()
        gdist.close_exchanges()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("xchg,batch", [("1", 8), ("0", 8), ("1", 64), ("1", 300)])   # NVLink peer exchange / NCCL all-gather of the
def test_two_gpus_match_one_gpu(xchg, batch, monkeypatch):                             # statistics; one-launch / two-launch kernels
    monkeypatch.setenv("GSSD_TEST_MULTI_B", str(batch))
    import torch.multiprocessing as mp
    from grouped_ssd_pytorch_b200 import dist as gdist
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, xchg)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    loc, conf, pri, tg = _inputs()
    ll, lc, gl, gc, m = _run(loc, conf, pri, tg, torch.device("cuda", 0))
    ref = np.array([ll.item(), lc.item()])
    for rank, tot, g_l, g_c, neg, pipe_ok in got:
        assert pipe_ok, "HostPipeline (begin / all-gather / finish) differs from MultiBoxLoss on rank %d" % rank
        sl = gdist.shard(loc.shape[0], rank, world)
        np.testing.assert_allclose(tot, ref, rtol=1e-6)
        assert np.array_equal(neg, m["neg"][sl].cpu().numpy())
        np.testing.assert_allclose(g_l, gl[sl].cpu().numpy(), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(g_c, gc[sl].cpu().numpy(), rtol=1e-6, atol=1e-12)
