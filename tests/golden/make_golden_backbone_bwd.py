"""Generate tests/golden/backbone_bwd.npz: what autograd computes through the reference's OWN backbone modules — the
[Conv2d(groups=4), BatchNorm2d, ReLU] triples conv3_2 .. conv5_3 that its vgg() builds (models/ssd_multiphase_custom_group.py:434-460),
taken from a model made by its build_ssd, cast to float64, in training mode — for the seeded cases of tests/cases.py: the pin of
source_block.PMConvLayer (tests/test_gpu_backbone_train.py).

Runs only in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_backbone_bwd.py

Per case the fixture keeps the output of the last ReLU and, of the input gradient and of every parameter gradient, a strided sample of at
most 1024 elements plus its sum and absolute sum; also the BatchNorm running statistics after the step.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/ssd_liverdet")
warnings.filterwarnings("ignore")

import torch  # noqa: E402

dcn = types.ModuleType("dcn_v2")
dcn._DCNv2 = type("_DCNv2", (), {"apply": staticmethod(lambda *a: (_ for _ in ()).throw(NotImplementedError()))})
sys.modules["dcn_v2"] = dcn
mpl = types.ModuleType("matplotlib")
mpl.use = lambda *a, **k: None
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")

from models.ssd_multiphase_custom_group import build_ssd  # noqa: E402  (reference)

import cases  # noqa: E402

torch.set_num_threads(4)
T = lambda a: torch.from_numpy(np.asarray(a)).double()


def sample(flat):
    flat = np.asarray(flat, np.float64).reshape(-1)
    step = max(1, flat.size // 1024)
    return flat[::step][:1024].astype(np.float32), np.array([flat.sum(), np.abs(flat).sum()], np.float64)


def run_case(tag, net):
    seed, first, n_triples, N, H, W = cases.BACKBONE_CASES[tag]
    x, prm, gout = cases.backbone_case(tag)
    mods = [net.vgg[k] for k in range(first, first + 3 * n_triples)]
    for t, p in enumerate(prm):
        conv, bn, relu = mods[3 * t:3 * t + 3]
        assert isinstance(conv, torch.nn.Conv2d) and conv.groups == 4 and isinstance(bn, torch.nn.BatchNorm2d) and isinstance(relu, torch.nn.ReLU)
        assert tuple(conv.weight.shape) == p["w"].shape, (tag, t, tuple(conv.weight.shape), p["w"].shape)
        with torch.no_grad():
            conv.weight.copy_(T(p["w"])); conv.bias.copy_(T(p["b"]))
            bn.weight.copy_(T(p["gamma"])); bn.bias.copy_(T(p["beta"]))
            bn.running_mean.zero_(); bn.running_var.fill_(1.0); bn.num_batches_tracked.zero_()
        bn.train()
        conv.zero_grad(); bn.zero_grad()
    xt = T(x).requires_grad_()
    h = xt
    for m in mods:                                             # GSSD:254-259 `x = self.vgg[k](x)`
        h = m(h)
    (h * T(gout)).sum().backward()
    out = {}
    out[tag + "/y_sample"], out[tag + "/y_sums"] = sample(h.detach().numpy())
    grads = {"x": xt.grad}
    for t in range(n_triples):
        conv, bn = mods[3 * t], mods[3 * t + 1]
        grads.update({"%d.conv_w" % t: conv.weight.grad, "%d.conv_b" % t: conv.bias.grad, "%d.bn_w" % t: bn.weight.grad, "%d.bn_b" % t: bn.bias.grad})
        out["%s/%d.running_mean" % (tag, t)] = bn.running_mean.numpy().astype(np.float32)
        out["%s/%d.running_var" % (tag, t)] = bn.running_var.numpy().astype(np.float32)
    for name, gr in grads.items():
        out[tag + "/" + name + "_sample"], out[tag + "/" + name + "_sums"] = sample(gr.numpy())
    return out


def main():
    # build_ssd(phase, size, num_classes, batch_norm, groups_vgg, groups_extra, feature_scale, use_fuseconv, ...)
    net = build_ssd('train', 300, 2, True, 4, 4, 1, True, False, False, 0, 1, False, False, 1).double()
    out = {}
    for tag in sorted(cases.BACKBONE_CASES):
        out.update(run_case(tag, net))
    path = os.path.join(HERE, "backbone_bwd.npz")
    np.savez_compressed(path, **out)
    print("backbone_bwd %8.1f KB, %d arrays" % (os.path.getsize(path) / 1024, len(out)))


if __name__ == "__main__":
    main()
