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
