"""Generate tests/golden/source_block.npz: the reference's OWN modules (built by its build_ssd) run on the
seeded source-block cases of tests/cases.py.

Runs only in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_block.py

For every case the fixture stores what the reference computes for one source of SSD.forward
(models/ssd_multiphase_custom_group.py): the statements at lines 258-259 / 300-301 (grouped conv, BN, ReLU),
281 (L2Norm), 290-297 / 317-323 / 365-369 (fuse conv, BN, ReLU) and 375-380 (heads, permute, flatten) are
executed verbatim on the model's modules after loading the case's seeded parameters into them.  Inputs and
parameters are regenerated in the tests from the seeds (tests/cases.py: block_case).
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

# import-time stubs for modules the reference model imports but GSSD (no DCN) never calls
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
T = torch.from_numpy


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
    # which of the reference's modules form this source (indices: SURVEY §3.1 / the module dump in DESIGN.md)
    if tag.startswith("s1"):
        conv_i = 30 if bn else 21
        gconv, gbn = net.vgg[conv_i], (net.vgg[conv_i + 1] if bn else None)
        fuse, bn_fuse, k = net.fuse_11, (net.bn_fuse_11 if bn else None), 0
    elif tag.startswith("s2"):
        gconv, gbn, fuse, bn_fuse, k = net.vgg[47], net.vgg[48], net.fuse_21, net.bn_fuse_21, 1
    else:
        gconv, gbn, fuse, bn_fuse, k = None, None, net.fuse_41, net.bn_fuse_41, 3
    if gconv is not None:
        load(gconv, prm, "gconv")
        if gbn is not None:
            load(gbn, prm, "bn", True)
    if l2:
        with torch.no_grad():
            net.L2Norm.weight.copy_(T(prm["l2norm_w"]))
    load(fuse, prm, "fuse")
    if bn_fuse is not None:
        load(bn_fuse, prm, "bn_fuse", True)
    load(net.loc[k], prm, "loc")
    load(net.conf[k], prm, "conf")
    with torch.no_grad():
        xt = T(x)
        if gconv is not None:
            xt = gconv(xt)                                                   # GSSD:258-259 `x = self.vgg[k](x)`
            if gbn is not None:
                xt = gbn(xt)
            xt = F.relu(xt, inplace=True)
        s = net.L2Norm(xt) if l2 else xt                                    # GSSD:281
        if bn_fuse is not None:
            s = F.relu(bn_fuse(fuse(s)), inplace=True)                      # GSSD:292 / 319 / 366
        else:
            s = F.relu(fuse(s), inplace=True)                               # GSSD:295
        loc = net.loc[k](s).permute(0, 2, 3, 1).contiguous()                # GSSD:376
        conf = net.conf[k](s).permute(0, 2, 3, 1).contiguous()              # GSSD:377
        out = {tag + "/x_out": xt.numpy(), tag + "/source": s.numpy(),
               tag + "/loc": loc.view(loc.size(0), -1).numpy(), tag + "/conf": conf.view(conf.size(0), -1).numpy()}
        if training:
            for nm, m in (("bn", gbn), ("bn_fuse", bn_fuse)):
                if m is not None:
                    out[tag + "/" + nm + "_running_mean"] = m.running_mean.numpy().copy()
                    out[tag + "/" + nm + "_running_var"] = m.running_var.numpy().copy()
    return out


def main():
    nets = {}
    for tag, c in cases.BLOCK_CASES.items():
        key = (c[6], c[10])
        if key not in nets:
            # build_ssd(phase, size, num_classes, batch_norm, groups_vgg, groups_extra, feature_scale, use_fuseconv, ...)
            nets[key] = build_ssd('train', 300, key[1], key[0], 4, 4, 1, True, False, False, 0, 1, False, False, 1)
    out = {}
    for tag in cases.BLOCK_CASES:
        out.update(run_case(tag, nets))
    path = os.path.join(HERE, "source_block.npz")
    np.savez_compressed(path, **{k: v.astype(np.float32) for k, v in out.items()})
    print("source_block %8.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
