"""A stand-in for the reference's GSSD model object (ssd_type gssd, SSD300, 4 phases), built with the same module
constructors and attribute names as models/ssd_multiphase_custom_group.py:40-139 (SSD.__init__), 432-460 (vgg),
463-490 (add_extras), 493-520 (multibox), 523-557 (tables, build_ssd) — the reference itself is not present on the GPU
box.  tests/golden/make_golden_model.py checks that its state_dict has the reference's keys and shapes."""
import numpy as np
import torch
import torch.nn as nn

BASE = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'C', 512, 512, 512, 'M', 512, 512, 512]
EXTRAS = [256, 'S', 512, 128, 'S', 256, 128, 256, 128, 256]
MBOX = [4, 6, 6, 6, 4, 4]


def vgg(batch_norm, groups=4, i=12):
    layers, c = [], i
    for v in BASE:
        if v == 'M':
            layers.append(nn.MaxPool2d(2, 2))
        elif v == 'C':
            layers.append(nn.MaxPool2d(2, 2, ceil_mode=True))
        else:
            layers += [nn.Conv2d(c, v, 3, padding=1, groups=groups)] + ([nn.BatchNorm2d(v)] if batch_norm else []) + [nn.ReLU(inplace=True)]
            c = v
    layers.append(nn.MaxPool2d(3, 1, 1))
    layers += [nn.Conv2d(512, 1024, 3, padding=6, dilation=6, groups=groups)] + ([nn.BatchNorm2d(1024)] if batch_norm else []) + [nn.ReLU(inplace=True)]
    layers += [nn.Conv2d(1024, 1024, 1, groups=groups)] + ([nn.BatchNorm2d(1024)] if batch_norm else []) + [nn.ReLU(inplace=True)]
    return layers


def add_extras(batch_norm, groups=4, i=1024):
    layers, c, flag = [], i, False
    for k, v in enumerate(EXTRAS):
        if c != 'S':
            if v == 'S':
                layers.append(nn.Conv2d(c, EXTRAS[k + 1], (1, 3)[flag], stride=2, padding=1, groups=groups))
                if batch_norm:
                    layers.append(nn.BatchNorm2d(EXTRAS[k + 1]))
            else:
                layers.append(nn.Conv2d(c, v, (1, 3)[flag], groups=groups))
                if batch_norm:
                    layers.append(nn.BatchNorm2d(v))
            flag = not flag
        c = v
    return layers


class StandInSSD(nn.Module):
    def __init__(self, phase, num_classes, batch_norm, priors):
        super().__init__()
        from grouped_ssd_pytorch_b200.layers import L2Norm
        self.phase, self.num_classes, self.batch_norm = phase, num_classes, batch_norm
        self.use_fuseconv, self.use_self_attention, self.use_self_attention_base, self.use_dcn = True, False, False, False
        self.priors = priors
        base, ext = vgg(batch_norm), add_extras(batch_norm)
        self.vgg = nn.ModuleList(base)
        self.L2Norm = L2Norm(512, 20)
        self.extras = nn.ModuleList(ext)
        src = [base[30 if batch_norm else 21], base[-3 if batch_norm else -2]] + list(ext[2::4] if batch_norm else ext[1::2])
        self.loc = nn.ModuleList([nn.Conv2d(s.out_channels, a * 4, 3, padding=1) for s, a in zip(src, MBOX)])
        self.conf = nn.ModuleList([nn.Conv2d(s.out_channels, a * num_classes, 3, padding=1) for s, a in zip(src, MBOX)])
        for name, c in (("11", 512), ("21", 1024), ("31", 512), ("41", 256), ("51", 256), ("61", 256)):
            setattr(self, "fuse_" + name, nn.Conv2d(c, c, 1))
            if batch_norm:
                setattr(self, "bn_fuse_" + name, nn.BatchNorm2d(c))
        self.fuse_list1 = nn.ModuleList([self.fuse_31, self.fuse_41, self.fuse_51, self.fuse_61])
        if batch_norm:
            self.bn_fuse_list1 = nn.ModuleList([self.bn_fuse_31, self.bn_fuse_41, self.bn_fuse_51, self.bn_fuse_61])


def seeded_state(state_dict, seed):
    """deterministic parameters for every entry of a state_dict, in key order (numpy stream: torch-version independent)"""
    r = np.random.RandomState(seed)
    out = {}
    for k, v in state_dict.items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=v.dtype)
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(r.uniform(0.5, 1.5, shape).astype(np.float32))
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy((r.randn(*shape) * 0.1).astype(np.float32))
        elif k == "L2Norm.weight":
            out[k] = torch.from_numpy((20 * r.uniform(0.8, 1.2, shape)).astype(np.float32))
        elif k.endswith(".weight") and len(shape) == 1:            # BatchNorm gamma
            out[k] = torch.from_numpy(r.uniform(0.8, 1.2, shape).astype(np.float32))
        elif k.endswith(".bias"):
            out[k] = torch.from_numpy((r.randn(*shape) * 0.05).astype(np.float32))
        else:                                                      # conv weight: He-normal keeps activations O(1)
            fan_in = int(np.prod(shape[1:]))
            out[k] = torch.from_numpy((r.randn(*shape) * np.sqrt(2.0 / fan_in)).astype(np.float32))
    return out


def seeded_input(seed, batch):
    r = np.random.RandomState(seed)
    return torch.from_numpy(r.uniform(0, 1, (batch, 12, 300, 300)).astype(np.float32))


def forward_torch(net, x):
    """the reference forward (ssd_multiphase_custom_group.py:217-396, ssd_type gssd, train-phase return) in plain torch"""
    import torch.nn.functional as F
    bn = net.batch_norm
    i43 = 33 if bn else 23
    sources, loc, conf = [], [], []
    for k in range(i43):
        x = net.vgg[k](x)
    s = net.L2Norm(x) if x.is_cuda else _l2norm_cpu(x, net.L2Norm)
    s = F.relu(net.bn_fuse_11(net.fuse_11(s)) if bn else net.fuse_11(s))
    sources.append(s)
    for k in range(i43, len(net.vgg)):
        x = net.vgg[k](x)
    sources.append(F.relu(net.bn_fuse_21(net.fuse_21(x)) if bn else net.fuse_21(x)))
    fc = 0
    for k, v in enumerate(net.extras):
        x = v(x)
        if bn:
            if k % 2 == 1:
                x = F.relu(x)
            src = k % 4 == 3
        else:
            x = F.relu(x)
            src = k % 2 == 1
        if src:
            z = net.fuse_list1[fc](x)
            sources.append(F.relu(net.bn_fuse_list1[fc](z) if bn else z))
            fc += 1
    for (xx, l, c) in zip(sources, net.loc, net.conf):
        loc.append(l(xx).permute(0, 2, 3, 1).contiguous())
        conf.append(c(xx).permute(0, 2, 3, 1).contiguous())
    loc = torch.cat([o.view(o.size(0), -1) for o in loc], 1)
    conf = torch.cat([o.view(o.size(0), -1) for o in conf], 1)
    return loc.view(loc.size(0), -1, 4), conf.view(conf.size(0), -1, net.num_classes)


def _l2norm_cpu(x, m):
    norm = x.pow(2).sum(dim=1, keepdim=True).sqrt() + m.eps       # l2norm.py:19-23 (the product L2Norm is CUDA-only)
    return m.weight.view(1, -1, 1, 1) * (x / norm)
