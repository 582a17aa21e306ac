"""Generate tests/golden/source_block_bwd.npz: gradients that autograd computes on the reference's OWN modules (built by its
build_ssd, cast to float64) for the seeded source-block cases of tests/cases.py — the pin of oracle.source_block's backward
(oracle for the conv backward planned in DESIGN.md §7).

Runs only in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_block_bwd.py

Per case: the statements of SSD.forward for one source (models/ssd_multiphase_custom_group.py:258-259 / 300-301, 281,
290-297 / 317-323 / 365-369, 375-380) run on the model's modules with the case's parameters, the scalar
sum(loc * d_loc) + sum(conf * d_conf) is back-propagated (d_loc, d_conf: cases.block_upstream) and the fixture keeps, of the
input gradient and of every parameter gradient, a strided sample of at most 1024 elements plus its sum and absolute sum.
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
import torch.nn.functional as F  # noqa: E402

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
TAGS = ("s1_train", "s2", "s4", "s1_nobn")


def load(mod, prm, name, bn=False):
    with torch.no_grad():
        mod.weight.copy_(T(prm[name + "_w"]))
        mod.bias.copy_(T(prm[name + "_b"]))
        if bn:
            mod.running_mean.copy_(T(prm[name + "_mean"]))
            mod.running_var.copy_(T(prm[name + "_var"]))


def run_case(tag, nets):
    x, prm, training = cases.block_case(tag)
    seed, N, C, H, W, gc, bn, l2, Cf, A, ncls, _ = cases.BLOCK_CASES[tag]
    net = nets[(bn, ncls)]
    net.train(training)
    net.zero_grad()
    if tag.startswith("s1"):
        conv_i = 30 if bn else 21
        gconv, gbn = net.vgg[conv_i], (net.vgg[conv_i + 1] if bn else None)
        fuse, bn_fuse, k = net.fuse_11, (net.bn_fuse_11 if bn else None), 0
    elif tag.startswith("s2"):
        gconv, gbn, fuse, bn_fuse, k = net.vgg[47], net.vgg[48], net.fuse_21, net.bn_fuse_21, 1
    else:
        gconv, gbn, fuse, bn_fuse, k = None, None, net.fuse_41, net.bn_fuse_41, 3
    mods = {}
    if gconv is not None:
        load(gconv, prm, "gconv"); mods["gconv"] = gconv
        if gbn is not None:
            load(gbn, prm, "bn", True); mods["bn"] = gbn
    if l2:
        with torch.no_grad():
            net.L2Norm.weight.copy_(T(prm["l2norm_w"]))
    load(fuse, prm, "fuse"); mods["fuse"] = fuse
    if bn_fuse is not None:
        load(bn_fuse, prm, "bn_fuse", True); mods["bn_fuse"] = bn_fuse
    load(net.loc[k], prm, "loc"); mods["loc"] = net.loc[k]
    load(net.conf[k], prm, "conf"); mods["conf"] = net.conf[k]
    xt = T(x).requires_grad_()
    h = xt
    if gconv is not None:
        h = gconv(h)                                                         # GSSD:258-259 `x = self.vgg[k](x)`
        if gbn is not None:
            h = gbn(h)
        h = F.relu(h)
    s = net.L2Norm(h) if l2 else h                                          # GSSD:281
    s = F.relu(bn_fuse(fuse(s))) if bn_fuse is not None else F.relu(fuse(s))   # GSSD:292 / 319 / 366, 295
    loc = net.loc[k](s).permute(0, 2, 3, 1).contiguous()                    # GSSD:376
    conf = net.conf[k](s).permute(0, 2, 3, 1).contiguous()                  # GSSD:377
    d_loc, d_conf = cases.block_upstream(tag, loc.view(N, -1).shape[1], conf.view(N, -1).shape[1])
    ((loc.view(N, -1) * T(d_loc)).sum() + (conf.view(N, -1) * T(d_conf)).sum()).backward()
    out = {}
    grads = {"x": xt.grad}
    grads.update({name + "_w": m.weight.grad for name, m in mods.items()})
    grads.update({name + "_b": m.bias.grad for name, m in mods.items()})
    if l2:
        grads["l2norm_w"] = net.L2Norm.weight.grad
    for name, gr in grads.items():
        flat = gr.numpy().reshape(-1)
        step = max(1, flat.size // 1024)
        out[tag + "/" + name + "_sample"] = flat[::step][:1024].astype(np.float32)
        out[tag + "/" + name + "_sums"] = np.array([flat.sum(), np.abs(flat).sum()], np.float64)
    return out


def main():
    nets = {}
    for tag in TAGS:
        c = cases.BLOCK_CASES[tag]
        key = (c[6], c[10])
        if key not in nets:
            # build_ssd(phase, size, num_classes, batch_norm, groups_vgg, groups_extra, feature_scale, use_fuseconv, ...)
            nets[key] = build_ssd('train', 300, key[1], key[0], 4, 4, 1, True, False, False, 0, 1, False, False, 1).double()
    out = {}
    for tag in TAGS:
        out.update(run_case(tag, nets))
    path = os.path.join(HERE, "source_block_bwd.npz")
    np.savez_compressed(path, **out)
    print("source_block_bwd %8.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
