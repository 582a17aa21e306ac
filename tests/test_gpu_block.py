"""GPU parity of the source block: the tcgen05/TMEM implicit-GEMM chain behind `SourceBlock`
against (1) the numpy oracle with bf16 operand emulation — the same arithmetic, tight tolerance — and (2) the
fixtures produced by the reference's own modules (tests/golden/source_block.npz) at the north-star tolerance for the
bf16 conv: 1e-2, taken relative to the tensor's scale (max |ref|), since single outputs pass through zero.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

import cases
from oracle import source_block as SB

pytestmark = pytest.mark.gpu

TIGHT, NORTH_STAR = 4e-3, 1e-2          # TIGHT = one bf16 ulp (2^-8) of the tensor's largest value


def rel(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-12))


def modules_from(tag, prm):
    """the reference's module constructors for this source (ssd_multiphase_custom_group.py:81-139, 434-520), loaded
    with the case's parameters"""
    from grouped_ssd_pytorch_b200.layers import L2Norm
    seed, N, C, H, W, gc, bn, l2, Cf, A, ncls, training = cases.BLOCK_CASES[tag]
    T = torch.from_numpy

    def load(m, name, is_bn=False):
        with torch.no_grad():
            m.weight.copy_(T(prm[name + "_w"])); m.bias.copy_(T(prm[name + "_b"]))
            if is_bn:
                m.running_mean.copy_(T(prm[name + "_mean"])); m.running_var.copy_(T(prm[name + "_var"]))
        return m

    gconv = gbn = l2m = bn_fuse = None
    c_mid = C
    if gc is not None:
        co, groups, k = gc
        gconv = load(nn.Conv2d(C, co, kernel_size=k, padding=(k - 1) // 2, groups=groups), "gconv")
        if bn:
            gbn = load(nn.BatchNorm2d(co), "bn", True)
        c_mid = co
    if l2:
        l2m = L2Norm(c_mid, 20)
        with torch.no_grad():
            l2m.weight.copy_(T(prm["l2norm_w"]))
    fuse = load(nn.Conv2d(c_mid, Cf, kernel_size=1), "fuse")
    if bn:
        bn_fuse = load(nn.BatchNorm2d(Cf), "bn_fuse", True)
    loc = load(nn.Conv2d(Cf, A * 4, kernel_size=3, padding=1), "loc")
    conf = load(nn.Conv2d(Cf, A * ncls, kernel_size=3, padding=1), "conf")
    mods = [m for m in (gconv, gbn, l2m, fuse, bn_fuse, loc, conf) if m is not None]
    for m in mods:
        m.cuda().train(training)
    return gconv, gbn, l2m, fuse, bn_fuse, loc, conf


def run_block(tag, prior_pad=(3, 5)):
    from grouped_ssd_pytorch_b200.layers import SourceBlock
    x, prm, training = cases.block_case(tag)
    seed, N, C, H, W, gc, bn, l2, Cf, A, ncls, _ = cases.BLOCK_CASES[tag]
    mods = modules_from(tag, prm)
    blk = SourceBlock(*mods, num_classes=ncls)
    n_pri = H * W * A
    P = prior_pad[0] + n_pri + prior_pad[1]                     # the slice lands inside a larger [B,P,*] tensor
    loc = torch.full((N, P, 4), 7.0, device="cuda")
    conf = torch.full((N, P, ncls), 7.0, device="cuda")
    x1, wrote = blk(torch.from_numpy(x).cuda(), loc, conf, prior_pad[0])
    torch.cuda.synchronize()
    assert wrote == n_pri
    # nothing outside the slice is touched
    assert bool((loc[:, :prior_pad[0]] == 7).all()) and bool((loc[:, prior_pad[0] + n_pri:] == 7).all())
    assert bool((conf[:, :prior_pad[0]] == 7).all()) and bool((conf[:, prior_pad[0] + n_pri:] == 7).all())
    out = dict(loc=loc[:, prior_pad[0]:prior_pad[0] + n_pri].reshape(N, -1).cpu().numpy(),
               conf=conf[:, prior_pad[0]:prior_pad[0] + n_pri].reshape(N, -1).cpu().numpy(),
               x_out=x1.to_nchw().cpu().numpy(), x1=x1)
    return out, (x, prm, training), mods


def test_layout_round_trip():
    from grouped_ssd_pytorch_b200.layers import PM
    r = np.random.RandomState(3)
    for shape in ((2, 64, 5, 7), (1, 512, 38, 38), (3, 192, 1, 1)):
        x = SB.bf16_round(r.randn(*shape).astype(np.float32))
        pm = PM.from_nchw(torch.from_numpy(x).cuda())
        n, c, h, w = shape
        grid = pm.data.float().view(n, h + 2, w + 2, c).cpu().numpy()
        assert np.array_equal(grid[:, 1:-1, 1:-1].transpose(0, 3, 1, 2), x)
        border = grid.copy(); border[:, 1:-1, 1:-1] = 0
        assert not border.any(), "the 1-pixel border must be zero"
        assert np.array_equal(pm.to_nchw().cpu().numpy(), x)


@pytest.mark.parametrize("tag", sorted(cases.BLOCK_CASES))
def test_source_block_matches_oracle_and_reference(tag):
    out, (x, prm, training), _ = run_block(tag)
    emu = SB.source_block(x, prm, training, emulate_bf16=True)
    g = cases.golden("source_block")
    for k in ("x_out", "loc", "conf"):
        assert rel(out[k], emu[k]) <= TIGHT, "%s/%s vs bf16-emulating oracle: %.3e" % (tag, k, rel(out[k], emu[k]))
        assert rel(out[k], g[tag + "/" + k]) <= NORTH_STAR, "%s/%s vs reference: %.3e" % (tag, k, rel(out[k], g[tag + "/" + k]))


def test_heads_wider_than_one_tile():
    """21 classes x 6 anchors = 150 head channels (the VOC configurations of the reference's data/config.py): the heads run as
    a padded plain convolution + the permute / flatten of GSSD:376-380"""
    from grouped_ssd_pytorch_b200.layers import SourceBlock
    torch.manual_seed(2)
    N, C, H, W, A, NC = 2, 256, 5, 7, 6, 21
    fuse, loc, conf = nn.Conv2d(C, C, 1).cuda(), nn.Conv2d(C, A * 4, 3, padding=1).cuda(), nn.Conv2d(C, A * NC, 3, padding=1).cuda()
    blk = SourceBlock(None, None, None, fuse, None, loc, conf, num_classes=NC)
    x = torch.relu(torch.randn(N, C, H, W, device="cuda"))
    P = H * W * A + 9
    lo, co = torch.full((N, P, 4), 7.0, device="cuda"), torch.full((N, P, NC), 7.0, device="cuda")
    with torch.no_grad():
        _, n = blk(x, lo, co, 4)
        s = torch.relu(fuse(x))
        want_l = loc(s).permute(0, 2, 3, 1).reshape(N, -1, 4)
        want_c = conf(s).permute(0, 2, 3, 1).reshape(N, -1, NC)
    assert n == H * W * A and bool((lo[:, :4] == 7).all()) and bool((co[:, 4 + n:] == 7).all())
    assert rel(lo[:, 4:4 + n].cpu().numpy(), want_l.cpu().numpy()) <= NORTH_STAR
    assert rel(co[:, 4:4 + n].cpu().numpy(), want_c.cpu().numpy()) <= NORTH_STAR


def test_border_of_the_block_output_is_zero():
    out, _, _ = run_block("s1")
    x1 = out["x1"]
    grid = x1.data.float().view(x1.n, x1.h + 2, x1.w + 2, x1.c)
    grid[:, 1:-1, 1:-1] = 0
    assert not bool(grid.any())


def test_train_mode_updates_running_statistics_like_batchnorm():
    out, (x, prm, training), mods = run_block("s1_train")
    g = cases.golden("source_block")
    gbn, bn_fuse = mods[1], mods[4]
    for nm, m in (("bn", gbn), ("bn_fuse", bn_fuse)):
        np.testing.assert_allclose(m.running_mean.cpu().numpy(), g["s1_train/%s_running_mean" % nm], rtol=0, atol=2e-3)
        np.testing.assert_allclose(m.running_var.cpu().numpy(), g["s1_train/%s_running_var" % nm], rtol=2e-2, atol=2e-3)
        assert int(m.num_batches_tracked) == 1


def test_full_size_source1_linearity_and_oracle_sample():
    """configs[1] shape for source 1 (batch 4 here): 38x38x512, groups 4.  Size-independent property: with biases
    zeroed and BN folded to identity the block's first conv is linear — conv(a*x) == a*conv(x) for a power of two
    (exact in bf16/fp32) — and a sample of output pixels matches a direct numpy evaluation."""
    from grouped_ssd_pytorch_b200.layers.modules.source_block import PM, _Conv, conv_igemm
    r = np.random.RandomState(9)
    N, C, H, W, groups = 4, 512, 38, 38, 4
    x = SB.bf16_round(np.maximum(r.randn(N, C, H, W), 0).astype(np.float32))
    w = SB.bf16_round((r.randn(C, C // groups, 3, 3) * np.sqrt(2.0 / (C // groups * 9))).astype(np.float32))
    conv = nn.Conv2d(C, C, 3, padding=1, groups=groups, bias=False).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(w))
    cv = _Conv(conv, groups, dev=torch.device("cuda"))
    pm = PM.from_nchw(torch.from_numpy(x).cuda())
    y1 = conv_igemm(pm, cv, relu=False).to_nchw()
    pm4 = PM.from_nchw(torch.from_numpy(x * 4).cuda())
    y4 = conv_igemm(pm4, cv, relu=False).to_nchw()
    assert torch.equal(y4, y1 * 4)
    y1 = y1.cpu().numpy()
    ref = SB.conv2d(x[:1], w, None, groups, 1)
    assert rel(y1[:1], SB.bf16_round(ref)) <= TIGHT
    # corners and edges of the other images (padding handling) against single-pixel dot products
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    for (n, yy, xx) in ((1, 0, 0), (2, 37, 37), (3, 0, 37), (3, 20, 0), (2, 37, 5)):
        for co in (0, 127, 128, 300, 511):
            g = co // (C // groups)
            patch = xp[n, g * (C // groups):(g + 1) * (C // groups), yy:yy + 3, xx:xx + 3].astype(np.float64)
            want = float((patch * w[co].astype(np.float64)).sum())
            assert abs(y1[n, co, yy, xx] - want) <= 2e-2 + 4e-3 * abs(want), (n, yy, xx, co, y1[n, co, yy, xx], want)


def test_train_mode_channel_statistics_at_full_size():
    """the per-channel sum / sum of squares the conv epilogue accumulates for train-mode BatchNorm (warp-transposed partial
    sums + red.add from ~400 tiles) against torch's own statistics of the same convolution, source-1 size, batch 4"""
    from grouped_ssd_pytorch_b200.layers.modules.source_block import PM, _Conv, conv_igemm
    r = np.random.RandomState(11)
    N, C, H, W, groups = 4, 512, 38, 38, 4
    x = torch.from_numpy(SB.bf16_round(np.maximum(r.randn(N, C, H, W), 0).astype(np.float32))).cuda()
    conv = nn.Conv2d(C, C, 3, padding=1, groups=groups).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(SB.bf16_round((r.randn(C, C // groups, 3, 3) * 0.03).astype(np.float32))))
        conv.bias.copy_(torch.from_numpy((r.randn(C) * 0.1).astype(np.float32)))
        torch.backends.cudnn.allow_tf32 = False
        ref = conv(x)
    cv = _Conv(conv, groups, dev=x.device)
    stats = torch.zeros(2 * C, device="cuda")
    y = conv_igemm(PM.from_nchw(x), cv, relu=False, shift=cv.bias, chan_sum=stats)
    n = N * H * W
    mean = stats[:C] / n
    var = stats[C:] / n - mean * mean
    np.testing.assert_allclose(mean.cpu().numpy(), ref.mean(dim=(0, 2, 3)).cpu().numpy(), rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(var.cpu().numpy(), ref.var(dim=(0, 2, 3), unbiased=False).cpu().numpy(), rtol=2e-3, atol=1e-5)
    # against cuDNN's accumulation order a value may round to the neighbouring bf16: one ulp of the top binade = 2^-7 of the max
    assert rel(y.to_nchw().cpu().numpy(), SB.bf16_round(ref.cpu().numpy())) <= 2.0 ** -7


# ---- backward of the chain: autograd through SourceBlock.forward_autograd ---------------------------------------------
def _grad_dict(xt, mods):
    gconv, gbn, l2m, fuse, bn_fuse, locm, confm = mods
    got = {"x": xt.grad}
    for name, m in (("gconv", gconv), ("bn", gbn), ("fuse", fuse), ("bn_fuse", bn_fuse), ("loc", locm), ("conf", confm)):
        if m is not None:
            got[name + "_w"], got[name + "_b"] = m.weight.grad, m.bias.grad
    if l2m is not None:
        got["l2norm_w"] = l2m.weight.grad
    return {k: v.detach().double().cpu().numpy() for k, v in got.items()}


def _backward_case(tag, with_xout=False):
    """-> (our gradients, torch-autograd gradients through the SAME modules with the ReLU masks of our forward, case)"""
    import torch.nn.functional as F
    from grouped_ssd_pytorch_b200.layers import SourceBlock
    x, prm, training = cases.block_case(tag)
    seed, N, C, H, W, gc, bn, l2, Cf, A, ncls, _ = cases.BLOCK_CASES[tag]
    mods = modules_from(tag, prm)
    gconv, gbn, l2m, fuse, bn_fuse, locm, confm = mods
    state = [(m, {k: v.clone() for k, v in m.state_dict().items()}) for m in mods if m is not None]
    blk = SourceBlock(*mods, num_classes=ncls)
    blk._debug_backward = {}
    xt = torch.from_numpy(x).cuda().requires_grad_()
    loc, conf, x_out = blk.forward_autograd(xt)
    d_loc, d_conf = cases.block_upstream(tag, loc[0].numel(), conf[0].numel())
    T = lambda a: torch.from_numpy(np.asarray(a, np.float32)).cuda()
    # with_xout: the block's second output (the post-ReLU grouped-conv map that continues down the backbone) takes part too
    w_out = None
    if with_xout and x_out is not None:
        w_out = T(np.random.RandomState(seed + 7).randn(*x_out.shape) * 0.05)
    total = (loc.reshape(N, -1) * T(d_loc)).sum() + (conf.reshape(N, -1) * T(d_conf)).sum()
    if w_out is not None:
        total = total + (x_out * w_out).sum()
    total.backward()
    torch.cuda.synchronize()
    got = _grad_dict(xt, mods)
    dbg = blk._debug_backward
    # the same chain by torch (fp32), ReLU replaced by OUR masks: a bf16 forward flips the sign of a few pre-activations that
    # are zero to within its rounding, and a flipped mask entry changes the gradients discontinuously — that is a property of
    # the forward's precision (covered by the forward tests), not of the backward under test
    for m, sd in state:
        m.load_state_dict(sd)                                    # running statistics as before our forward
        m.zero_grad()
    m1 = (dbg["y1"] > 0).float() if gconv is not None else None
    m2 = (dbg["z2"] > 0).float()
    xr = torch.from_numpy(x).cuda().requires_grad_()
    h = xr
    if gconv is not None:
        h = gconv(h)
        if gbn is not None:
            h = gbn(h)
        h = h * m1
    s_ = l2m(h) if l2m is not None else h
    z = fuse(s_)
    if bn_fuse is not None:
        z = bn_fuse(z)
    z = z * m2
    lo = locm(z).permute(0, 2, 3, 1).reshape(N, -1)
    co = confm(z).permute(0, 2, 3, 1).reshape(N, -1)
    total = (lo * T(d_loc)).sum() + (co * T(d_conf)).sum()
    if w_out is not None:
        total = total + (h * w_out).sum()
    total.backward()
    ref = _grad_dict(xr, mods)
    flips = 0
    with torch.no_grad():
        hh = xr
        if gconv is not None:
            hh = gconv(hh)
            hh = gbn(hh) if gbn is not None else hh
            flips += int(((hh > 0).float() != m1).sum())
    return got, ref, flips, (x, prm, training)


@pytest.mark.parametrize("tag", ["s1_train", "s2", "s1_nobn"])
def test_source_block_backward_with_the_gradient_of_its_second_output(tag):
    """the post-ReLU grouped-conv map continues down the backbone (GSSD:300-301): its upstream gradient joins the gradient that
    comes back through L2Norm / fuse before the ReLU mask and the BatchNorm backward"""
    got, ref, _, _ = _backward_case(tag, with_xout=True)
    peers = max(np.abs(v).max() for v in ref.values())
    for name in sorted(ref):
        a, r = got[name].reshape(-1), ref[name].reshape(-1)
        if np.abs(r).max() < 1e-5 * peers:
            continue
        err = np.abs(a - r).max() / np.abs(r).max()
        assert err <= 2e-2, "%s/%s: %.3e of the tensor's scale" % (tag, name, err)


@pytest.mark.parametrize("tag", ["s1_train", "s2", "s4", "s1_nobn"])
def test_source_block_backward(tag):
    """Every gradient autograd produces through the chain — input, conv filters and biases, BatchNorm weight / bias in training
    AND eval mode, L2Norm weight — from the tcgen05 backward: data gradients on the forward kernel, weight gradients on
    gssd_conv_wgrad, BN / ReLU / L2Norm backward on PM rows.
      (1) against torch autograd (fp32) through the same modules with the ReLU masks of our forward: max error <= 1e-2 of each
          tensor's scale, the north-star tolerance of the bf16 conv block (2e-2 for the gradients that pass through all five
          bf16 tensors of the backward);
      (2) against the reference's own autograd in float64 (tests/golden/source_block_bwd.npz): relative L2 error of the sampled
          entries.  Max-norm cannot be used there: on these 50-pixel maps one flipped ReLU mask entry (a pre-activation that is
          zero to within bf16 rounding; a handful per case) moves a whole filter row by more than 10 %."""
    g = cases.golden("source_block_bwd")
    got, ref, flips, _ = _backward_case(tag)
    names = sorted(k[len(tag) + 1:-len("_sample")] for k in g.files if k.startswith(tag + "/") and k.endswith("_sample"))
    assert set(names) == set(got) == set(ref), (names, sorted(got))
    peers = max(np.abs(ref[n]).max() for n in names)
    worst, l2err, bad = {}, {}, []
    for name in names:
        a, r = got[name].reshape(-1), ref[name].reshape(-1)
        vanishing = np.abs(r).max() < 1e-5 * peers                          # conv bias in front of a training-mode BN
        scale = max(np.abs(r).max(), 1e-6 * peers)
        err = np.abs(a - r).max() / scale
        worst[name] = err
        deep = name in ("x", "gconv_w", "gconv_b", "bn_w", "bn_b", "l2norm_w")
        if vanishing:
            if np.abs(a).max() > 1e-2 * peers:
                bad.append(name + " (should vanish)")
            continue
        if err > (2e-2 if deep else 1e-2):
            bad.append(name)
        step = max(1, a.size // 1024)
        gs = g[tag + "/" + name + "_sample"].astype(np.float64)
        l2err[name] = float(np.linalg.norm(a[::step][:1024] - gs) / max(np.linalg.norm(gs), 1e-30))
        if l2err[name] > 1e-1:
            bad.append(name + " (vs the reference's float64 autograd)")
    print("test_source_block_backward[%s]: max error / scale vs torch with our masks: %s" % (tag, {k: "%.1e" % v for k, v in worst.items()}))
    print("    relative L2 error vs the reference's float64 autograd (%d flipped ReLU masks in the first stage): %s" % (
        flips, {k: "%.1e" % v for k, v in l2err.items()}))
    assert not bad, "%s: gradients off by more than the tolerance: %s" % (tag, bad)


def test_source_block_backward_full_size_linearity_and_oracle_sample():
    """configs[1] size (batch 4 x 38x38 x 512, training-mode BN + L2Norm): the backward is linear in the upstream gradient, and
    sampled entries of the filter gradients agree with dot products taken by torch from the saved activations"""
    from grouped_ssd_pytorch_b200.layers import L2Norm, SourceBlock
    torch.manual_seed(5)
    N, C, H, A, NC = 4, 512, 38, 4, 2
    gconv, gbn = nn.Conv2d(C, C, 3, padding=1, groups=4).cuda(), nn.BatchNorm2d(C).cuda()
    fuse, bnf = nn.Conv2d(C, C, 1).cuda(), nn.BatchNorm2d(C).cuda()
    l2 = L2Norm(C, 20).cuda()
    loc, conf = nn.Conv2d(C, A * 4, 3, padding=1).cuda(), nn.Conv2d(C, A * NC, 3, padding=1).cuda()
    blk = SourceBlock(gconv, gbn, l2, fuse, bnf, loc, conf, num_classes=NC)
    mods = [gconv, gbn, l2, fuse, bnf, loc, conf]
    x = torch.relu(torch.randn(N, C, H, H, device="cuda"))

    def grads(up_l, up_c):
        for m in mods:
            m.zero_grad()
        xt = x.clone().requires_grad_()
        lo, co, xo = blk.forward_autograd(xt)
        ((lo * up_l).sum() + (co * up_c).sum()).backward()
        return [xt.grad.clone()] + [p.grad.clone() for m in mods for p in m.parameters()]

    P = H * H * A
    u1, u2 = torch.randn(N, P, 4, device="cuda"), torch.randn(N, P, NC, device="cuda")
    v1, v2 = torch.randn(N, P, 4, device="cuda"), torch.randn(N, P, NC, device="cuda")
    ga, gb, gab = grads(u1, u2), grads(v1, v2), grads(u1 + v1, u2 + v2)
    peers_ab = max(float(t.abs().max()) for t in gab)
    for a, b, ab in zip(ga, gb, gab):
        scale = max(float(ab.abs().max()), 1e-3 * peers_ab)                 # (a conv bias in front of a training-mode BN has no gradient)
        assert float((a + b - ab).abs().max()) <= 5e-2 * scale                # bf16 rounding of the intermediate gradients, three passes
    # against torch autograd on the same modules (fp32 cuDNN), same inputs, with the ReLU masks of our forward (see
    # test_source_block_backward): the reference semantics at full size
    blk._debug_backward = {}
    grads(u1, u2)
    m1, m2 = (blk._debug_backward["y1"] > 0).float(), (blk._debug_backward["z2"] > 0).float()
    blk._debug_backward = None
    for m in mods:
        m.zero_grad()
    xt = x.clone().requires_grad_()
    h = gbn(gconv(xt)) * m1
    s = bnf(fuse(l2(h))) * m2
    lo = loc(s).permute(0, 2, 3, 1).reshape(N, -1, 4)
    co = conf(s).permute(0, 2, 3, 1).reshape(N, -1, NC)
    ((lo * u1).sum() + (co * u2).sum()).backward()
    ref = [xt.grad.clone()] + [p.grad.clone() for m in mods for p in m.parameters()]
    names = ["x"] + [mn + "." + pn for m, mn in zip(mods, ["gconv", "bn", "l2norm", "fuse", "bn_fuse", "loc", "conf"])
                     for pn, _ in m.named_parameters()]
    peers = max(float(r.abs().max()) for r in ref)
    worst = {}
    for n, a, r in zip(names, ga, ref):
        if float(r.abs().max()) < 1e-5 * peers:
            continue                                                         # conv bias in front of a training-mode BN
        worst[n] = float((a - r).abs().max()) / float(r.abs().max())
    print("full-size backward vs torch autograd with our masks, max error / scale:", {k: "%.1e" % v for k, v in worst.items()})
    # the maximum runs over up to 3 M entries here and the split-K additions of the weight gradient are unordered (the 50-pixel
    # cases above are held to 1e-2 / 2e-2): 4e-2; measured 4e-3 .. 2.4e-2 over several runs
    assert all(v <= 4e-2 for v in worst.values()), worst
