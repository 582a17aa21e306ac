"""The one-launch MultiBoxLoss (csrc/fused.cu: gssd_mbox_loss_fused) against the two-stage path (gssd_mbox_match +
gssd_mbox_loss) and against the oracle, through the C ABI: same masks, same gradients bit for bit, losses to 1e-6
(the per-CTA partial sums are added in a different order), for every cluster size / CTA width the kernel can be launched with."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
from grouped_ssd_pytorch_b200 import _lib, synthetic as syn
from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _both(B, pname, gmax, C, ratio=3, seed=0, conf=None, tg=None):
    lib = _lib.require_cuda()
    dev = torch.device("cuda:0")
    pri_np = cases.priors(pname)
    P = pri_np.shape[0]
    r = syn.rng(500 + B + seed)
    tg = tg if tg is not None else syn.targets(r, B, 1, gmax)
    if C > 2:
        for t in tg:
            t[:, 4] = r.randint(0, C - 1, size=t.shape[0])
    loc_np = syn.loc(r, B, P)
    conf_np = conf if conf is not None else syn.conf_logits(r, B, P, C)
    pri, loc, cf = (torch.from_numpy(x).to(dev) for x in (pri_np, loc_np, conf_np))
    gt, gt_off, sum_g, g_max = pack_target_list([torch.from_numpy(t) for t in tg], dev)
    st = _lib.stream()
    wsb = lib.gssd_workspace_bytes(_lib.WS_LOSS, B, P, C, sum_g, 0)

    def bufs():
        return dict(losses=torch.empty(2, device=dev), gl=torch.full_like(loc, 7.0), gc=torch.full_like(cf, 7.0),
                    pos=torch.empty(B, P, dtype=torch.uint8, device=dev), neg=torch.empty(B, P, dtype=torch.uint8, device=dev),
                    ws=torch.empty(wsb, dtype=torch.uint8, device=dev))

    assert lib.gssd_mbox_fused_supported(B, P, C, g_max) == 1, "shape expected to have a one-launch form"
    f = bufs()
    npos = torch.empty(B, dtype=torch.int32, device=dev)
    state = torch.zeros(int(lib.gssd_fused_state_bytes()), dtype=torch.uint8, device=dev)
    for _ in range(3):                                           # several launches on the same state: every launch is a new epoch
        _lib.check(lib.gssd_mbox_loss_fused(loc.data_ptr(), cf.data_ptr(), pri.data_ptr(), B, P, C, gt.data_ptr(), gt_off.data_ptr(),
                                            sum_g, g_max, 0.5, ratio, 0.1, 0.2, state.data_ptr(), None, f["losses"].data_ptr(),
                                            f["gl"].data_ptr(), f["gc"].data_ptr(), f["pos"].data_ptr(), f["neg"].data_ptr(),
                                            npos.data_ptr(), f["ws"].data_ptr(), wsb, st), "gssd_mbox_loss_fused")
    torch.cuda.synchronize()
    assert int(state[:4].view(torch.int32)) == 3, "the epoch advances once per launch"
    t = bufs()
    tags = torch.empty(B, P, dtype=torch.int16, device=dev)
    stats = torch.empty(16 + 4 * B, dtype=torch.uint8, device=dev)
    _lib.check(lib.gssd_mbox_match(pri.data_ptr(), P, cf.data_ptr(), C, gt.data_ptr(), gt_off.data_ptr(), B, sum_g, g_max, 0.5,
                                   tags.data_ptr(), stats.data_ptr(), st), "gssd_mbox_match")
    _lib.check(lib.gssd_mbox_loss(loc.data_ptr(), cf.data_ptr(), pri.data_ptr(), B, P, C, gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max,
                                  tags.data_ptr(), stats.data_ptr(), None, 0, ratio, 0.1, 0.2, t["losses"].data_ptr(), t["gl"].data_ptr(),
                                  t["gc"].data_ptr(), t["pos"].data_ptr(), t["neg"].data_ptr(), t["ws"].data_ptr(), wsb, st), "gssd_mbox_loss")
    torch.cuda.synchronize()
    t["npos"] = stats[16:].view(torch.int32)
    f["npos"] = npos
    return f, t, (loc_np, conf_np, pri_np, tg)


def _compare(f, t):
    assert torch.equal(f["npos"], t["npos"])
    assert torch.equal(f["pos"], t["pos"]) and torch.equal(f["neg"], t["neg"])
    assert torch.equal(f["gl"], t["gl"]), "grad_loc differs from the two-stage path"
    assert torch.equal(f["gc"], t["gc"]), "grad_conf differs from the two-stage path"
    np.testing.assert_allclose(f["losses"].cpu().numpy(), t["losses"].cpu().numpy(), rtol=1e-6)


@pytest.mark.parametrize("B,pname,gmax,C", [(32, "v2", 5, 2), (1, "v2", 5, 2), (3, "v2", 3, 2), (74, "v2", 5, 2), (148, "v2", 2, 2),
                                            (8, "v2_512", 32, 2), (16, "v2_512", 32, 2), (5, "small", 3, 2), (4, "v2", 6, 4),
                                            (2, "v2_custom_512", 100, 2)])
def test_fused_equals_two_stage(B, pname, gmax, C):
    f, t, (loc, conf, pri, tg) = _both(B, pname, gmax, C)
    _compare(f, t)
    if B <= 8:                                                   # and the oracle, where it is quick
        o = O.multibox_loss(loc, conf, pri, tg, 0.5, 3, cases.VAR)
        np.testing.assert_array_equal(f["pos"].cpu().numpy(), o["pos"])
        np.testing.assert_array_equal(f["npos"].cpu().numpy(), o["num_pos"])
        np.testing.assert_allclose(f["losses"].cpu().numpy(), [o["loss_l"], o["loss_c"]], rtol=1e-5)


@pytest.mark.parametrize("pname,step", [("small", 0.25), ("v2", 0.002), ("v2_512", 0.001)])
def test_fused_key_ties(pname, step):
    """heavily tied mining keys: blocks of 8 equal keys, every second key exactly 0, every key identical — the select then needs
    its later passes (more than 256 composites share 11 / 22 / 32 leading bits; the prior index decides)"""
    P = cases.priors(pname).shape[0]
    conf = np.zeros((3, P, 2), np.float32)
    conf[0, :, 1] = np.repeat(np.arange(P // 8 + 1), 8)[:P] * step
    conf[1] = syn.conf_logits(syn.rng(5), 1, P, 2)[0]
    conf[1, ::2] = np.array([40.0, -40.0], np.float32)
    conf[2, :, 1] = 1.0
    f, t, (loc, conf, pri, tg) = _both(3, pname, 3, 2, conf=conf, tg=syn.targets(syn.rng(77), 3, 2, 3))
    _compare(f, t)
    o = O.multibox_loss(loc, conf, pri, tg, 0.5, 3, cases.VAR)
    np.testing.assert_array_equal(f["neg"].cpu().numpy(), o["neg"])
    # ratio large enough that num_neg is clamped to P - 1 (multibox_loss.py:105) and ratio 0 (no negatives at all)
    for ratio in (1000, 0):
        f, t, _ = _both(3, pname, 3, 2, ratio=ratio, conf=conf, tg=syn.targets(syn.rng(77), 3, 2, 3))
        _compare(f, t)


_FORCED = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import test_gpu_fused as T
for (B, pname, gmax, C) in %(shapes)r:
    f, t, _ = T._both(B, pname, gmax, C)
    T._compare(f, t)
f, t, _ = T._both(3, "v2", 3, 2, conf=None)
print("FORCED-OK")
"""


@pytest.mark.parametrize("S,NT", [(1, 256), (1, 1024), (2, 512), (4, 256), (8, 256), (4, 512), (2, 1024), (4, 1024)])
def test_fused_every_launch_shape(S, NT):
    """GSSD_FUSED_S / GSSD_FUSED_NT force the cluster size and CTA width (read once per process, hence the subprocess)"""
    shapes = [(16, "v2", 5, 2), (3, "v2_512" if S > 1 else "small", 32, 2), (2, "v2", 4, 3)]
    env = dict(os.environ, GSSD_FUSED_S=str(S), GSSD_FUSED_NT=str(NT))
    r = subprocess.run([sys.executable, "-c", _FORCED % dict(root=ROOT, shapes=shapes)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "FORCED-OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_fused_module_path_and_graph_replay():
    """MultiBoxLoss takes the one-launch path for the training batch; replaying it from a CUDA graph gives the same numbers"""
    from grouped_ssd_pytorch_b200.layers import MultiBoxLoss
    loc, conf, pri, tg, C, ratio = cases.loss_case("a")
    crit = MultiBoxLoss(C, 0.5, True, 0, True, ratio, 0.5, False, True)
    l = torch.from_numpy(loc).cuda().requires_grad_(); c = torch.from_numpy(conf).cuda().requires_grad_()
    pr = torch.from_numpy(pri).cuda()
    tgt = pack_target_list([torch.from_numpy(t) for t in tg], torch.device("cuda:0"))
    n0 = _lib.launch_count()
    ll, lc = crit((l, c, pr), tgt)
    (ll + lc).backward()
    assert _lib.launch_count() - n0 == 2, "one launch for the loss + one for the backward rescale"
    ref = (ll.item(), lc.item(), l.grad.clone(), c.grad.clone())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            l.grad = None; c.grad = None
            ll, lc = crit((l, c, pr), tgt); (ll + lc).backward()
    torch.cuda.current_stream().wait_stream(s)
    del ll, lc                                                   # no autograd graph of an earlier stream alive during the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l.grad = None; c.grad = None
    with torch.cuda.graph(g):
        ll, lc = crit((l, c, pr), tgt)
        (ll + lc).backward()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert (ll.item(), lc.item()) == ref[:2]
    assert torch.equal(l.grad, ref[2]) and torch.equal(c.grad, ref[3])
