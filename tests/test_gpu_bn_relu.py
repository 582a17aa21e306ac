"""Training-mode BatchNorm2d + ReLU of the NCHW backbone layers on gssd_bn_relu_nchw_fwd / _bwd (layers/modules/bn_relu.py):
against the numpy oracle (oracle/source_block.py: batch_norm / batch_norm_backward, pinned against the reference's own modules)
and against torch's nn.BatchNorm2d + F.relu in fp32 — outputs, input / weight / bias gradients, running statistics,
num_batches_tracked — on vectorised and scalar plane sizes, with and without the ReLU, and through `run_layers` on a slice of a
backbone-like nn.ModuleList."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from grouped_ssd_pytorch_b200 import _lib
from grouped_ssd_pytorch_b200.layers.modules.bn_relu import bn_relu, max_pool, run_layers, takes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, ref):
    a, ref = a.detach().double().cpu().numpy(), np.asarray(ref.detach().double().cpu().numpy() if isinstance(ref, torch.Tensor) else ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.mark.parametrize("shape,relu", [((4, 16, 30, 30), True), ((2, 8, 5, 7), True), ((3, 4, 1, 1), True), ((2, 12, 38, 38), False),
                                         ((8, 64, 150, 150), True)])
def test_bn_relu_vs_torch_and_oracle(shape, relu):
    torch.manual_seed(sum(shape))
    N, C, H, W = shape
    x = (torch.randn(shape, device=DEV) * 1.7 + 0.4).requires_grad_(True)
    bn_a, bn_b = nn.BatchNorm2d(C).to(DEV).train(), nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5); bn_a.bias.normal_(0, 0.3); bn_a.running_mean.normal_(); bn_a.running_var.uniform_(0.5, 2)
    bn_b.load_state_dict(bn_a.state_dict())
    gout = torch.randn(shape, device=DEV)
    n0 = _lib.launch_count()
    y = bn_relu(x, bn_a, relu=relu)
    y.backward(gout)
    assert _lib.launch_count() == n0 + 4, "two kernels forward, two backward"
    gx, x.grad = x.grad, None
    yr = bn_b(x)
    yr = F.relu(yr) if relu else yr
    yr.backward(gout)
    errs = dict(y=rel(y, yr), dx=rel(gx, x.grad), dw=rel(bn_a.weight.grad, bn_b.weight.grad), db=rel(bn_a.bias.grad, bn_b.bias.grad),
                rm=rel(bn_a.running_mean, bn_b.running_mean), rv=rel(bn_a.running_var, bn_b.running_var))
    print(shape, relu, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 2e-5, errs
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 1
    if N * C * H * W <= 100000:                                            # the oracle in float64
        from oracle import source_block as SB
        xn = x.detach().cpu().numpy().astype(np.float64)
        g, b = bn_b.weight.detach().cpu().numpy().astype(np.float64), bn_b.bias.detach().cpu().numpy().astype(np.float64)
        mean, var = xn.mean((0, 2, 3)), xn.var((0, 2, 3))
        yo = SB.batch_norm(xn, g, b, mean, var, bn_b.eps, True)[0].astype(np.float64)
        go = gout.cpu().numpy().astype(np.float64)
        if relu:
            go = go * (yo > 0)
            yo = np.maximum(yo, 0)
        dxo = SB.batch_norm_backward(xn, g, mean, var, bn_b.eps, True, go)
        dxo = dxo[0] if isinstance(dxo, tuple) else dxo
        assert rel(y, torch.from_numpy(yo)) <= 2e-5 and rel(gx, torch.from_numpy(dxo)) <= 5e-5


def test_run_layers_fuses_pairs_and_leaves_the_rest():
    torch.manual_seed(3)
    torch.backends.cudnn.allow_tf32 = False                                 # the convolutions are cuDNN's on both sides: keep them fp32
    mods = nn.ModuleList([nn.Conv2d(12, 16, 3, padding=1, groups=4), nn.BatchNorm2d(16), nn.ReLU(inplace=True), nn.MaxPool2d(2, 2),
                          nn.Conv2d(16, 32, 3, padding=1, groups=4), nn.BatchNorm2d(32), nn.ReLU(inplace=True)]).to(DEV).train()
    with torch.no_grad():
        mods[0].bias.normal_(0, 2); mods[4].bias.normal_(0, 2)              # biases that matter for the running means
    import copy
    ref = copy.deepcopy(nn.Sequential(*mods))                               # torch's own modules on the same parameters
    x = torch.randn(3, 12, 20, 20, device=DEV, requires_grad=True)
    n0 = _lib.launch_count()
    y = run_layers(mods, x)
    y.square().sum().backward()
    assert _lib.launch_count() == n0 + 9                                    # two fused pairs + the pool's backward
    gx, x.grad = x.grad, None
    # the yardstick is torch's own modules in float64: with large convolution biases cuDNN's fp32 batch statistics lose digits
    # (the mean dominates the variance), while the bias-free fused path does not see the bias at all
    ref32 = copy.deepcopy(ref)
    ref = ref.double()
    xd = x.detach().double().requires_grad_(True)
    yr = ref(xd)
    yr.square().sum().backward()
    y32 = ref32(x.detach())
    print("run_layers vs torch float64: %.1e   (torch float32 modules vs float64: %.1e)" % (rel(y, yr), rel(y32, yr)))
    assert rel(y, yr) <= 2e-5 and rel(gx, xd.grad) <= 1e-4
    for (n, p), (_, q) in zip(mods.named_parameters(), ref.named_parameters()):
        if n in ("0.bias", "4.bias"):                                       # a convolution's bias in front of a training-mode BatchNorm has
            continue                                                        # no gradient (BN removes the mean): rounding noise on both sides
        assert rel(p.grad, q.grad) <= 1e-4, n
    for i in (1, 5):                                                        # the convolutions ran without their bias: it must still reach the
        assert rel(mods[i].running_mean, ref[i].running_mean) <= 1e-5       # running mean (and nothing else)
        assert rel(mods[i].running_var, ref[i].running_var) <= 1e-5
    assert float(mods[0].bias.grad.abs().max()) == 0.0 and float(mods[4].bias.grad.abs().max()) == 0.0
    mods.eval()
    assert not takes(x, mods[1])                                            # evaluation mode stays torch's
    with torch.no_grad():
        assert rel(run_layers(mods, x), ref.eval()(x.double())) <= 1e-3     # (running statistics after one step agree)


def test_bn_relu_argument_errors():
    bn = nn.BatchNorm2d(4).to(DEV).eval()
    with pytest.raises(NotImplementedError):
        bn_relu(torch.randn(1, 4, 3, 3, device=DEV), bn)
    lib = _lib.load()
    assert lib.gssd_bn_relu_nchw_fwd(None, None, None, 1, 4, 9, 1e-5, 1, None, None, None, None, 0.1, None, None, None) == _lib.ERR_ARG


def test_bn_relu_bandwidth_at_the_backbone_size():
    """conv1_x of the batch-32 training step: [32, 64, 300, 300] fp32 (737 MB).  Reported, and held to twice cuDNN's speed."""
    x = torch.randn(32, 64, 300, 300, device=DEV, requires_grad=True)
    gout = torch.randn_like(x)
    bn = nn.BatchNorm2d(64).to(DEV).train()

    def timed(fn, iters=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def ours():
        x.grad = None
        bn_relu(x, bn).backward(gout)

    def ref():
        x.grad = None
        F.relu(bn(x)).backward(gout)
    t_o, t_r = timed(ours), timed(ref)
    gb = x.numel() * 4 * 8 / 1e9                                            # 3 passes forward, 5 backward
    print("BN + ReLU forward + backward at [32, 64, 300, 300]: ours %.2f ms (%.0f GB/s of the 8 algorithmic passes), torch / cuDNN %.2f ms" % (t_o, gb / t_o * 1e3, t_r))
    assert t_o < t_r


@pytest.mark.parametrize("shape,k,s,p,ceil", [((2, 8, 30, 30), 2, 2, 0, False), ((2, 4, 75, 75), 2, 2, 0, True), ((3, 6, 19, 19), 3, 1, 1, False),
                                               ((1, 5, 10, 11), 3, 2, 1, True), ((2, 3, 7, 7), 3, 3, 0, False), ((4, 64, 150, 150), 2, 2, 0, False)])
def test_max_pool_backward_is_torchs(shape, k, s, p, ceil):
    """the pools of the backbone (ssd_multiphase_custom_group.py:437-446: 2x2/2, 2x2/2 ceil_mode, 3x3/1 pad 1) and a few others; ReLU-like
    inputs with many exact ties (zeros): the argmax is torch's own (the forward is its kernel), the gradient must be bit-equal"""
    torch.manual_seed(k * 10 + s)
    x = torch.relu(torch.randn(shape, device=DEV)).requires_grad_(True)
    m = nn.MaxPool2d(k, s, p, ceil_mode=ceil)
    gout = torch.randn(m(x).shape, device=DEV)
    n0 = _lib.launch_count()
    max_pool(x, m).backward(gout)
    assert _lib.launch_count() == n0 + 1
    gx, x.grad = x.grad, None
    m(x).backward(gout)
    assert torch.equal(gx, x.grad) or float((gx - x.grad).abs().max()) <= 1e-6 * float(x.grad.abs().max())    # (overlapping windows: summation order)
    with torch.no_grad():
        assert torch.equal(max_pool(x.detach(), m), m(x.detach()))


@pytest.mark.parametrize("shape,relu", [((4, 16, 30, 30), True), ((2, 64, 5, 7), True), ((2, 12, 19, 19), False), ((3, 1024, 3, 3), True),
                                         ((8, 64, 150, 150), True)])
def test_bn_relu_channels_last_vs_torch(shape, relu):
    """the channels-last form (gssd_bn_relu_nhwc_*): same numbers, output and input gradient stay channels-last"""
    torch.manual_seed(sum(shape) + 1)
    N, C, H, W = shape
    x = (torch.randn(shape, device=DEV) * 1.3 - 0.2).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    bn_a, bn_b = nn.BatchNorm2d(C).to(DEV).train(), nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5); bn_a.bias.normal_(0, 0.3)
    bn_b.load_state_dict(bn_a.state_dict())
    gout = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
    y = bn_relu(x, bn_a, relu=relu)
    nhwc = C % 4 == 0 and 256 % (C // 4) == 0                          # else the NCHW kernels serve a contiguous copy
    assert y.is_contiguous(memory_format=torch.channels_last) == nhwc or (H == 1 and W == 1)
    y.backward(gout)
    gx, x.grad = x.grad, None
    yr = bn_b(x)
    yr = F.relu(yr) if relu else yr
    yr.backward(gout)
    errs = dict(y=rel(y, yr), dx=rel(gx, x.grad), dw=rel(bn_a.weight.grad, bn_b.weight.grad), db=rel(bn_a.bias.grad, bn_b.bias.grad),
                rm=rel(bn_a.running_mean, bn_b.running_mean), rv=rel(bn_a.running_var, bn_b.running_var))
    print(shape, relu, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) <= 2e-5, errs


@pytest.mark.parametrize("shape,k,s,p,ceil", [((2, 8, 30, 30), 2, 2, 0, False), ((2, 4, 75, 75), 2, 2, 0, True), ((3, 8, 19, 19), 3, 1, 1, False),
                                               ((1, 12, 10, 11), 3, 2, 1, True), ((2, 16, 7, 7), 2, 2, 0, False), ((4, 64, 150, 150), 2, 2, 0, False)])
def test_max_pool_backward_channels_last(shape, k, s, p, ceil):
    torch.manual_seed(k * 10 + s + 1)
    x = torch.relu(torch.randn(shape, device=DEV)).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    m = nn.MaxPool2d(k, s, p, ceil_mode=ceil)
    gout = torch.randn(m(x).shape, device=DEV).contiguous(memory_format=torch.channels_last)
    out = max_pool(x, m)
    out.backward(gout)
    gx, x.grad = x.grad, None
    assert gx.is_contiguous(memory_format=torch.channels_last)
    m(x).backward(gout)
    assert torch.equal(gx, x.grad) or float((gx - x.grad).abs().max()) <= 1e-6 * float(x.grad.abs().max())
