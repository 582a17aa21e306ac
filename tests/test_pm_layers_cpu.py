"""Host logic of the grouped backbone convolutions under autograd (source_block.PMConvLayer): the autograd
plumbing between the PM tensors, the group merging of the weight gradient (`_wgrad_any`), the rotated filter of the data
gradient and the run detection of `run_layers(tc=...)` — with the kernels replaced by torch emulations of their contracts
(include/gssd.h), so that it runs without a GPU.  The kernels themselves are checked on the GPU by
tests/test_gpu_backbone_train.py."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from grouped_ssd_pytorch_b200.layers.modules import source_block as SB
from grouped_ssd_pytorch_b200.layers.modules import bn_relu as BR


class _EmuConv(object):
    """stands in for source_block._Conv: keeps the fp32 filter instead of the packed bf16 one"""

    def __init__(self, conv, groups, in_scale=None, extra=None, dev=None, pad_out=None):
        assert in_scale is None and extra is None and pad_out is None
        self.weight = conv.weight.detach().float().to(torch.bfloat16).float()      # the kernel's operands are bf16
        self.groups = groups
        self.c_out, self.cg = self.weight.shape[0], self.weight.shape[1]
        self.c_in = self.cg * groups
        self.taps = self.weight.shape[2] * self.weight.shape[3]
        self.bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(self.c_out)
        self.scale, self.shift = None, self.bias


def _pm_from_nchw(x):
    x = x.detach().float()
    n, c, h, w = x.shape
    return SB.PM(F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(-1, c).to(torch.bfloat16).contiguous(), n, c, h, w)


def _pm_to_nchw(self):
    g = self.data.float().view(self.n, self.h + 2, self.w + 2, self.c)
    return g[:, 1:-1, 1:-1].permute(0, 3, 1, 2).contiguous()


def _conv_igemm(x, cv, relu, y=True, row_ss_in=None, l2_eps=1e-10, row_ss_out=None, chan_sum=None, scale=None, shift=None, head=None):
    assert head is None and row_ss_in is None and row_ss_out is None and scale is None
    k = int(round(cv.taps ** 0.5))
    out = F.conv2d(_pm_to_nchw(x), cv.weight, shift, padding=k // 2, groups=cv.groups)
    if relu:
        out = out.relu()
    if chan_sum is not None:                                       # fp32 sums of the raw output over the interior pixels
        chan_sum[:cv.c_out] += out.sum(dim=(0, 2, 3))
        chan_sum[cv.c_out:] += out.square().sum(dim=(0, 2, 3))
    return _pm_from_nchw(out)


def _bn_train(y, bn, stats, want_ss, out=None):
    assert not want_ss and out is not None
    cnt = y.n * y.h * y.w
    mean = stats[:y.c] / cnt
    var = (stats[y.c:] / cnt - mean * mean).clamp_min(0)
    v = _pm_to_nchw(y)
    z = (v - mean.view(1, -1, 1, 1)) * torch.rsqrt(var + bn.eps).view(1, -1, 1, 1)
    if bn.affine:
        z = z * bn.weight.detach().view(1, -1, 1, 1) + bn.bias.detach().view(1, -1, 1, 1)
    out.data.copy_(_pm_from_nchw(z.relu()).data)
    with torch.no_grad():
        bn.num_batches_tracked += 1
        bn.running_mean.mul_(1 - bn.momentum).add_(mean, alpha=bn.momentum)
        bn.running_var.mul_(1 - bn.momentum).add_(var * cnt / (cnt - 1), alpha=bn.momentum)
    return None


def _bn_relu_bwd(dy, y, yraw, add, stats, gamma, bn_eps, ss_l2, l2_eps, ss_out, l2_eps_out, ebn=None):
    assert add is None and ss_l2 is None and ss_out is None and ebn is None and stats is not None
    c, cnt = dy.c, dy.n * dy.h * dy.w
    mean = stats[:c] / cnt
    rstd = torch.rsqrt((stats[c:] / cnt - mean * mean).clamp_min(0) + bn_eps)
    g = _pm_to_nchw(dy) * (_pm_to_nchw(y) > 0)
    xh = (_pm_to_nchw(yraw) - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
    sg, sgx = g.sum(dim=(0, 2, 3)), (g * xh).sum(dim=(0, 2, 3))
    gam = gamma if gamma is not None else torch.ones(c)
    dx = (gam * rstd).view(1, -1, 1, 1) * (g - (sg / cnt).view(1, -1, 1, 1) - xh * (sgx / cnt).view(1, -1, 1, 1))
    return _pm_from_nchw(dx), torch.cat([sg, sgx, dx.sum(dim=(0, 2, 3))])


def _wgrad(dy, x, c_out, groups, taps):
    """the contract of gssd_conv_wgrad: channels per group in multiples of 128, like the kernel"""
    k = int(round(taps ** 0.5))
    cg, ng = x.c // groups, c_out // groups
    assert cg % 128 == 0 and (groups == 1 or ng % 128 == 0), "gssd_conv_wgrad would return GSSD_ERR_LIMIT"
    return torch.nn.grad.conv2d_weight(_pm_to_nchw(x), (c_out, cg, k, k), _pm_to_nchw(dy)[:, :c_out], padding=k // 2, groups=groups)


@pytest.fixture
def emulated(monkeypatch):
    monkeypatch.setattr(SB, "_Conv", _EmuConv)
    monkeypatch.setattr(SB.PM, "from_nchw", staticmethod(_pm_from_nchw))
    monkeypatch.setattr(SB.PM, "to_nchw", _pm_to_nchw)
    monkeypatch.setattr(SB.PM, "empty", staticmethod(lambda n, c, h, w, dev: SB.PM(torch.zeros((n * (h + 2) * (w + 2), c), dtype=torch.bfloat16), n, c, h, w)))
    monkeypatch.setattr(SB, "conv_igemm", _conv_igemm)
    monkeypatch.setattr(SB.SourceBlock, "_bn_train", staticmethod(_bn_train))
    monkeypatch.setattr(SB, "_bn_relu_bwd", _bn_relu_bwd)
    monkeypatch.setattr(SB, "_wgrad", _wgrad)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: __import__("contextlib").nullcontext())
    real = SB.PMConvLayer.takes

    def takes(conv, bn, relu, x=None, width=None):                  # as the real one, minus "x is a CUDA tensor"
        ok = real(conv, bn, relu, None, width)
        return ok and (x is None or (x.dim() == 4 and x.shape[1] == conv.in_channels and real(conv, bn, relu, None, x.shape[3])))
    monkeypatch.setattr(SB.PMConvLayer, "takes", staticmethod(takes))


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max())


def rel2(a, b):
    """relative L2 error: a ReLU mask entry that flips under the bf16 rounding of the forward moves single gradient entries by
    far more than the rounding itself (see tests/test_gpu_block.py), the L2 norm is robust against the handful that do"""
    return float((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm())


def _stack(specs, seed):
    torch.manual_seed(seed)
    mods = []
    for cin, cout, g, k in specs:
        conv = nn.Conv2d(cin, cout, k, padding=k // 2, groups=g)
        bn = nn.BatchNorm2d(cout)
        with torch.no_grad():
            conv.weight.copy_((torch.randn_like(conv.weight) * 0.05).to(torch.bfloat16).float())
            conv.bias.normal_(0, 0.1)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
        mods += [conv, bn, nn.ReLU(inplace=True)]
    return nn.ModuleList(mods).train()


@pytest.mark.parametrize("specs", [
    [(256, 256, 4, 3), (256, 256, 4, 3)],                          # conv3_2 / conv3_3: 64 channels per group, merged in pairs
    [(256, 512, 4, 3), (512, 512, 4, 3)],                          # conv4_1 / conv4_2: 64 -> 128 and 128 -> 128 per group
    [(128, 256, 2, 1)],                                            # a 1x1 triple
])
def test_pm_layers_match_torch_autograd(emulated, specs):
    import copy
    mods = _stack(specs, 5)
    ref = copy.deepcopy(mods).double()
    x = (torch.randn(2, specs[0][0], 6, 5)).to(torch.bfloat16).float().requires_grad_()
    n0 = SB.PMConvLayer.calls
    cache = {}
    y = BR.run_layers(mods, x, tc=cache)
    assert SB.PMConvLayer.calls == n0 + len(specs) and len(cache) == len(specs)
    w_out = torch.randn_like(y)
    (y * w_out).sum().backward()
    xd = x.detach().double().requires_grad_()
    h = xd
    for m in ref:
        h = m(h)
    (h * w_out.double()).sum().backward()
    assert y.shape == h.shape and rel(y, h) <= 2e-2
    assert rel2(x.grad, xd.grad) <= 8e-2          # plumbing errors (a wrong group block, a missing rotation) are O(1)
    for (name, p), (_, q) in zip(mods.named_parameters(), ref.named_parameters()):
        if name.endswith("bias") and isinstance(mods[int(name.split(".")[0])], nn.Conv2d):
            assert float(p.grad.abs().max()) <= 2e-2 * float(w_out.abs().sum())     # vanishes in front of a training-mode BatchNorm
            continue
        assert p.grad is not None and p.grad.shape == q.grad.shape, name
        assert rel2(p.grad, q.grad) <= 8e-2, (name, rel2(p.grad, q.grad))
    for i in range(1, len(mods), 3):                               # running statistics as nn.BatchNorm2d
        assert rel(mods[i].running_mean, ref[i].running_mean) <= 1e-2 and rel(mods[i].running_var, ref[i].running_var) <= 1e-2
        assert int(mods[i].num_batches_tracked) == 1


def test_wgrad_any_drops_the_cross_group_blocks(emulated):
    r = np.random.RandomState(2)
    n, h, w, groups, cg, ng = 2, 5, 4, 4, 64, 64
    x = torch.from_numpy(r.randn(n, groups * cg, h, w).astype(np.float32))
    dy = torch.from_numpy(r.randn(n, groups * ng, h, w).astype(np.float32))
    got = SB._wgrad_any(_pm_from_nchw(dy), _pm_from_nchw(x), groups * ng, groups, 9)
    want = torch.nn.grad.conv2d_weight(_pm_to_nchw(_pm_from_nchw(x)), (groups * ng, cg, 3, 3), _pm_to_nchw(_pm_from_nchw(dy)), padding=1, groups=groups)
    assert got.shape == want.shape and rel(got, want) <= 1e-5
    # 32 channels per group: four groups as one
    got = SB._wgrad_any(_pm_from_nchw(dy[:, :128]), _pm_from_nchw(x[:, :128]), 128, 4, 9)
    want = torch.nn.grad.conv2d_weight(_pm_to_nchw(_pm_from_nchw(x[:, :128])), (128, 32, 3, 3), _pm_to_nchw(_pm_from_nchw(dy[:, :128])), padding=1, groups=4)
    assert got.shape == want.shape and rel(got, want) <= 1e-5
    with pytest.raises(NotImplementedError):                       # three groups of 64 do not pair up
        SB._wgrad_any(_pm_from_nchw(dy[:, :192]), _pm_from_nchw(x[:, :192]), 192, 3, 9)


def test_run_detection_stops_at_foreign_layers(emulated):
    mods = nn.ModuleList([nn.Conv2d(128, 256, 3, padding=1, groups=4), nn.BatchNorm2d(256), nn.ReLU(),      # 32 per group: torch
                          nn.Conv2d(256, 256, 3, padding=1, groups=4), nn.BatchNorm2d(256), nn.ReLU(),
                          nn.Conv2d(256, 256, 3, padding=1, groups=4), nn.BatchNorm2d(256), nn.ReLU(),
                          nn.MaxPool2d(2, 2),
                          nn.Conv2d(256, 512, 3, padding=1, groups=4), nn.BatchNorm2d(512), nn.ReLU(),
                          nn.Conv2d(512, 512, 3, padding=6, dilation=6, groups=4), nn.BatchNorm2d(512), nn.ReLU()]).train()
    x = torch.zeros(1, 256, 8, 8)
    assert SB.pm_layers_at(mods, 0, len(mods), torch.zeros(1, 128, 8, 8)) == []
    assert [c for c, _ in SB.pm_layers_at(mods, 3, len(mods), x)] == [mods[3], mods[6]]
    assert [c for c, _ in SB.pm_layers_at(mods, 3, 8, x)] == [mods[3]]                  # the slice ends inside the second triple
    assert [c for c, _ in SB.pm_layers_at(mods, 10, len(mods), torch.zeros(1, 256, 4, 4))] == [mods[10]]   # dilated conv6 stays torch
    assert SB.pm_layers_at(mods, 10, len(mods), torch.zeros(1, 256, 4, 100)) == []      # too wide for the slab ring at 128 per group
    mods.eval()
    assert SB.pm_layers_at(mods, 3, len(mods), x) == []                                 # evaluation mode: BackboneRun's business
