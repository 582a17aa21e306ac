"""SourceBlock — the per-source chain of the reference's SSD.forward
(models/ssd_multiphase_custom_group.py:258-297, 300-325, 329-372 and the heads at 375-380):

    grouped conv (groups=4) -> BN -> ReLU -> [L2Norm] -> fuse_X1 (1x1) -> bn_fuse_X1 -> ReLU
        -> loc.k / conf.k (3x3) -> permute(0,2,3,1) -> flatten -> concat

run as calls of libgssd_b200.so's tcgen05/TMEM implicit-GEMM convolution (`gssd_conv_igemm`).  The block does
not own parameters: it is built FROM the reference's own modules (`net.vgg[30]`, `net.vgg[31]`,
`net.L2Norm`, `net.fuse_11`, `net.bn_fuse_11`, `net.loc[0]`, `net.conf[0]` ...), so state-dict names and
checkpoints are the reference's.  Weights are re-packed to bf16 whenever a parameter changes.

Forward only (inference and the forward half of a training step; the backward is SURVEY §8f rank 1).
No CPU fallback: without a CUDA device every call raises.
"""
import ctypes as C

import torch

from ... import _lib


class PM(object):
    """pixel-major padded activation: bf16 [n*(h+2)*(w+2), c] with a zero 1-pixel border (include/gssd.h)."""
    __slots__ = ("data", "n", "c", "h", "w")

    def __init__(self, data, n, c, h, w):
        self.data, self.n, self.c, self.h, self.w = data, n, c, h, w

    @property
    def rows(self):
        return self.n * (self.h + 2) * (self.w + 2)

    @staticmethod
    def empty(n, c, h, w, dev):
        return PM(torch.empty((n * (h + 2) * (w + 2), c), dtype=torch.bfloat16, device=dev), n, c, h, w)

    @staticmethod
    def from_nchw(x):
        lib = _lib.require_cuda()
        dev = _lib.device_of(x)
        with torch.cuda.device(dev):
            xc = _lib.f32(x, dev)
            n, c, h, w = xc.shape
            out = PM.empty(n, c, h, w, dev)
            _lib.check(lib.gssd_nchw_to_pm(xc.data_ptr(), n, c, h, w, out.data.data_ptr(), _lib.stream()), "gssd_nchw_to_pm")
        return out

    def to_nchw(self):
        lib = _lib.require_cuda()
        with torch.cuda.device(self.data.device):
            y = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.data.device)
            _lib.check(lib.gssd_pm_to_nchw(self.data.data_ptr(), self.n, self.c, self.h, self.w, y.data_ptr(), _lib.stream()),
                       "gssd_pm_to_nchw")
        return y


def maxpool_pm(x, pool):
    """nn.MaxPool2d on a PM tensor (gssd_maxpool_pm)."""
    lib = _lib.require_cuda()
    one = lambda v: v if isinstance(v, int) else v[0]
    k, st, pd = one(pool.kernel_size), one(pool.stride), one(pool.padding)
    if one(pool.dilation) != 1:
        raise NotImplementedError("dilated pooling")
    oh, ow = C.c_int(), C.c_int()
    _lib.check(lib.gssd_maxpool_pm(None, x.n, x.c, x.h, x.w, k, st, pd, int(bool(pool.ceil_mode)), None, C.byref(oh), C.byref(ow), None))
    out = PM.empty(x.n, x.c, oh.value, ow.value, x.data.device)
    with torch.cuda.device(x.data.device):
        _lib.check(lib.gssd_maxpool_pm(x.data.data_ptr(), x.n, x.c, x.h, x.w, k, st, pd, int(bool(pool.ceil_mode)),
                                       out.data.data_ptr(), None, None, _lib.stream()), "gssd_maxpool_pm")
    return out


def _conv_eligible(conv):
    """stride-1 1x1 / 3x3 'same' convolutions whose per-group channel counts are multiples of 64 run on gssd_conv_igemm"""
    import torch.nn as nn
    if not isinstance(conv, nn.Conv2d):
        return False
    kh, kw = conv.kernel_size
    return ((kh, kw) in ((1, 1), (3, 3)) and conv.stride == (1, 1) and conv.dilation == (1, 1)
            and conv.padding == ((kh - 1) // 2, (kw - 1) // 2)
            and (conv.in_channels // conv.groups) % 64 == 0 and (conv.out_channels // conv.groups) % 64 == 0)


class BackboneRun(object):
    """A stretch of the model's `vgg` ModuleList executed in PM/bf16: every [Conv2d, (BatchNorm2d), ReLU] triple whose
    conv fits the tcgen05 kernel runs as one launch with BN folded and ReLU fused (eval mode), MaxPool2d runs on PM, and
    anything else (conv6: dilation 6) round-trips through torch.  SURVEY §8f rank 1, forward half."""

    def __init__(self, modules):
        self.modules = list(modules)
        self._packed = {}
        self._too_wide = set()

    def _conv(self, i, conv, bn, dev):
        key = (_versions(conv, bn), str(dev))
        hit = self._packed.get(i)
        if hit is None or hit[0] != key:
            with torch.no_grad(), torch.cuda.device(dev):
                cv = _Conv(conv, conv.groups, dev=dev)
                if bn is not None:
                    cv.fold_bn(bn)
            hit = (key, cv)
            self._packed[i] = hit
        return hit[1]

    def __call__(self, x, start, stop):
        """x: PM or NCHW tensor; runs modules[start:stop]; returns PM"""
        import torch.nn as nn
        mods = self.modules
        i = start
        while i < stop:
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < stop else None
            if _conv_eligible(m):
                bn = nxt if isinstance(nxt, nn.BatchNorm2d) else None
                j = i + (2 if bn is not None else 1)
                relu = j < stop and isinstance(mods[j], nn.ReLU)
                if bn is not None and bn.training:
                    raise NotImplementedError("BackboneRun folds BatchNorm: eval mode only")
                if i in self._too_wide:                              # the slab of this feature-map width does not fit: torch
                    if isinstance(x, PM):
                        x = x.to_nchw()
                    x = m(x)
                    i += 1
                    continue
                xin = x if isinstance(x, PM) else PM.from_nchw(x)
                cv = self._conv(i, m, bn, xin.data.device)
                try:
                    x = conv_igemm(xin, cv, relu=relu, scale=cv.scale, shift=cv.shift)
                except RuntimeError as e:
                    if "GSSD_MAX" not in str(e) and "limits" not in str(e):
                        raise
                    self._too_wide.add(i)
                    continue
                i = j + (1 if relu else 0)
            elif isinstance(m, nn.MaxPool2d) and isinstance(x, PM):
                x = maxpool_pm(x, m)
                i += 1
            else:
                if isinstance(x, PM):
                    x = x.to_nchw()
                x = m(x)
                i += 1
        return x if isinstance(x, PM) else PM.from_nchw(x)


def _versions(*mods):
    v = []
    for m in mods:
        if m is None:
            continue
        for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False)):
            v.append((t.data_ptr(), t._version))
    return tuple(v)


class _Conv(object):
    """one packed convolution: bf16 weights [c_out(+pad), taps*c_in/groups] + fp32 epilogue vectors"""

    def __init__(self, conv, groups, in_scale=None, extra=None, dev=None):
        lib = _lib.require_cuda()
        ws = [conv.weight] + ([extra.weight] if extra is not None else [])
        bs = [conv.bias] + ([extra.bias] if extra is not None else [])
        w = torch.cat([_lib.f32(t, dev) for t in ws], 0)
        kh, kw = conv.kernel_size
        if (kh, kw) not in ((1, 1), (3, 3)) or conv.stride != (1, 1) or conv.dilation != (1, 1) or \
                conv.padding != ((kh - 1) // 2, (kw - 1) // 2):
            raise NotImplementedError("gssd_conv_igemm takes 1x1 and 3x3 / stride 1 / pad 1 convolutions, got %r" % (conv,))
        self.taps = kh * kw
        self.groups = groups
        self.c_out, self.cg = w.shape[0], w.shape[1]
        self.c_in = self.cg * groups
        rows = self.c_out if extra is None else -(-self.c_out // 32) * 32      # head weights: whole 32-row TMA boxes
        self.w = torch.zeros((rows, self.taps * self.cg), dtype=torch.bfloat16, device=dev)
        sc = None if in_scale is None else _lib.f32(in_scale, dev)
        _lib.check(lib.gssd_conv_pack_weights(w.data_ptr(), self.c_out, self.cg, groups, self.taps, _lib.ptr(sc),
                                              self.w.data_ptr(), _lib.stream()), "gssd_conv_pack_weights")
        if all(b is None for b in bs):
            self.bias = torch.zeros((self.c_out,), dtype=torch.float32, device=dev)
        else:
            self.bias = torch.cat([_lib.f32(b, dev) if b is not None else torch.zeros((t.shape[0],), device=dev)
                                   for b, t in zip(bs, ws)], 0)
        self.scale, self.shift = None, self.bias

    def fold_bn(self, bn):
        """eval-mode BatchNorm folded into the epilogue: y = acc*s + (beta + (bias - mean)*s), s = gamma/sqrt(var+eps)"""
        dev = self.bias.device
        s = _lib.f32(bn.weight, dev) / torch.sqrt(_lib.f32(bn.running_var, dev) + bn.eps)
        self.scale = s.contiguous()
        self.shift = (_lib.f32(bn.bias, dev) + (self.bias - _lib.f32(bn.running_mean, dev)) * s).contiguous()


def dgrad_weight(weight, groups):
    """Filter with which the data gradient of a stride-1 'same' convolution is itself such a convolution of dY:
    d/dx conv2d(x, w, padding=k//2, groups=g) . dY == conv2d(dY, dgrad_weight(w, g), padding=k//2, groups=g) —
    rotated by 180 degrees, in / out channels swapped inside each group ([Cout, Cin/g, k, k] -> [Cin, Cout/g, k, k]).
    Host-side half of the backward planned in DESIGN.md §7: packed with `gssd_conv_pack_weights`, it puts dgrad on the
    same tcgen05 kernel as the forward."""
    cout, cin_g, kh, kw = weight.shape
    wt = weight.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2).flip(3, 4)
    return wt.reshape(groups * cin_g, cout // groups, kh, kw).contiguous()


def conv_igemm(x, cv, relu, y=True, row_ss_in=None, l2_eps=1e-10, row_ss_out=None, chan_sum=None, scale=None, shift=None,
               head=None):
    """One `gssd_conv_igemm` call.  x: PM; cv: _Conv; head = (loc, conf, n_anchor, n_cls, prior_off, n_priors)."""
    lib = _lib.require_cuda()
    if x.c != cv.c_in:
        raise ValueError("conv expects %d input channels, got %d" % (cv.c_in, x.c))
    d = _lib.ConvDesc()
    d.n_img, d.height, d.width = x.n, x.h, x.w
    d.c_in, d.c_out, d.groups, d.taps, d.relu = cv.c_in, cv.c_out, cv.groups, cv.taps, 1 if relu else 0
    d.x, d.w = x.data.data_ptr(), cv.w.data_ptr()
    d.scale, d.shift = _lib.ptr(scale), _lib.ptr(shift)
    d.row_ss_in, d.l2_eps = _lib.ptr(row_ss_in), float(l2_eps)
    out = None
    if head is None and y:
        out = PM.empty(x.n, cv.c_out, x.h, x.w, x.data.device)
        d.y = out.data.data_ptr()
    d.row_ss_out, d.chan_sum = _lib.ptr(row_ss_out), _lib.ptr(chan_sum)
    if head is not None:
        loc, conf, d.n_anchor, d.n_cls, d.prior_off, d.n_priors = head
        d.loc, d.conf = loc.data_ptr(), conf.data_ptr()
    with torch.cuda.device(x.data.device):
        _lib.check(lib.gssd_conv_igemm(C.byref(d), _lib.stream()), "gssd_conv_igemm")
    return out


class SourceBlock(object):
    """gconv/bn may be None (sources 3-6: the input is already the post-ReLU feature map); l2norm only for source 1.
    `momentum` updates of the BatchNorm running statistics follow nn.BatchNorm2d in training mode."""

    def __init__(self, gconv, bn, l2norm, fuse, bn_fuse, loc, conf, num_classes):
        self.gconv, self.bn, self.l2norm, self.fuse, self.bn_fuse, self.loc, self.conf = gconv, bn, l2norm, fuse, bn_fuse, loc, conf
        self.num_classes = num_classes
        self.n_anchor = loc.out_channels // 4
        if conf.out_channels != self.n_anchor * num_classes:
            raise ValueError("conf head has %d channels, expected %d anchors x %d classes" % (conf.out_channels, self.n_anchor, num_classes))
        self._packed, self._key = None, None

    def _pack(self, dev):
        mods = (self.gconv, self.bn, self.l2norm, self.fuse, self.bn_fuse, self.loc, self.conf)
        training = bool((self.bn is not None and self.bn.training) or (self.bn_fuse is not None and self.bn_fuse.training))
        key = (_versions(*mods), training, str(dev))
        if self._key == key:
            return self._packed
        with torch.no_grad(), torch.cuda.device(dev):
            g = _Conv(self.gconv, self.gconv.groups, dev=dev) if self.gconv is not None else None
            f = _Conv(self.fuse, 1, in_scale=self.l2norm.weight if self.l2norm is not None else None, dev=dev)
            h = _Conv(self.loc, 1, extra=self.conf, dev=dev)
            if not training:
                if g is not None and self.bn is not None:
                    g.fold_bn(self.bn)
                if self.bn_fuse is not None:
                    f.fold_bn(self.bn_fuse)
        self._packed, self._key = (g, f, h, training), key
        return self._packed

    @staticmethod
    def _bn_train(y, bn, stats, want_ss):
        """batch-statistics BatchNorm + ReLU in place on the raw conv output; running stats as nn.BatchNorm2d"""
        lib = _lib.load()
        dev = y.data.device
        ss = torch.empty((y.rows,), dtype=torch.float32, device=dev) if want_ss else None
        mv = torch.empty((2 * y.c,), dtype=torch.float32, device=dev) if bn.track_running_stats else None
        with torch.cuda.device(dev):
            _lib.check(lib.gssd_bn_act_pm(y.data.data_ptr(), y.n, y.c, y.h, y.w, stats.data_ptr(),
                                          _lib.ptr(_lib.f32(bn.weight, dev)) if bn.affine else None,
                                          _lib.ptr(_lib.f32(bn.bias, dev)) if bn.affine else None,
                                          float(bn.eps), 1, _lib.ptr(ss), _lib.ptr(mv), _lib.stream()), "gssd_bn_act_pm")
        if mv is not None:
            with torch.no_grad():
                bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - mom).add_(mv[:y.c].to(bn.running_mean.device), alpha=mom)
                bn.running_var.mul_(1 - mom).add_(mv[y.c:].to(bn.running_var.device), alpha=mom)
        return ss

    def forward(self, x, loc_out, conf_out, prior_off):
        """x: NCHW fp32 tensor or PM.  Writes this source's slice of loc_out[B,P,4] / conf_out[B,P,C] (fp32, CUDA) at
        prior offset `prior_off`; returns (x_out PM = the block's post-ReLU grouped-conv output, n_priors_written)."""
        if not isinstance(x, PM):
            x = PM.from_nchw(x)
        dev = x.data.device
        g, f, h, training = self._pack(dev)
        P = loc_out.shape[1]
        want_l2 = self.l2norm is not None
        with torch.no_grad():
            ss = torch.empty((x.rows,), dtype=torch.float32, device=dev) if want_l2 else None
            if g is not None:
                if training and self.bn is not None:
                    stats = torch.zeros((2 * g.c_out,), dtype=torch.float32, device=dev)
                    x1 = conv_igemm(x, g, relu=False, shift=g.bias, chan_sum=stats)
                    ss = self._bn_train(x1, self.bn, stats, want_l2)
                else:
                    x1 = conv_igemm(x, g, relu=True, scale=g.scale, shift=g.shift, row_ss_out=ss)
            else:
                x1 = x
                if want_l2:
                    raise NotImplementedError("L2Norm without the grouped conv in front")
            eps = self.l2norm.eps if want_l2 else 0.0
            if training and self.bn_fuse is not None:
                stats = torch.zeros((2 * f.c_out,), dtype=torch.float32, device=dev)
                src = conv_igemm(x1, f, relu=False, shift=f.bias, chan_sum=stats, row_ss_in=ss, l2_eps=eps)
                self._bn_train(src, self.bn_fuse, stats, False)
            else:
                src = conv_igemm(x1, f, relu=True, scale=f.scale, shift=f.shift, row_ss_in=ss, l2_eps=eps)
            conv_igemm(src, h, relu=False, shift=h.bias,
                       head=(loc_out, conf_out, self.n_anchor, self.num_classes, int(prior_off), P))
        return x1, x.h * x.w * self.n_anchor

    __call__ = forward


# ---- the model's forward with the source blocks swapped in ---------------------------------------------------
def build_source_blocks(net):
    """SourceBlocks for the six sources of a reference `SSD` (ssd_type gssd: no self-attention, no DCN), built from
    its own modules.  Module indices follow multibox()'s `vgg_source` (ssd_multiphase_custom_group.py:499-502):
    conv4_3 = vgg[30] (batch_norm) / vgg[21]; conv7 = vgg[-3] / vgg[-2]."""
    if getattr(net, "use_self_attention", False) or getattr(net, "use_self_attention_base", False) or getattr(net, "use_dcn", False):
        raise NotImplementedError("gssd_forward covers ssd_type 'gssd'; GSSD++ (self-attention / DCN) keeps the reference forward")
    if not getattr(net, "use_fuseconv", True):
        raise NotImplementedError("the source blocks expect use_fuseconv=True (the GSSD configuration)")
    bn = bool(net.batch_norm)
    vgg = net.vgg
    i43 = 30 if bn else 21
    i7 = len(vgg) - (3 if bn else 2)
    nc = net.num_classes
    blocks = [SourceBlock(vgg[i43], vgg[i43 + 1] if bn else None, net.L2Norm, net.fuse_11, net.bn_fuse_11 if bn else None,
                          net.loc[0], net.conf[0], nc),
              SourceBlock(vgg[i7], vgg[i7 + 1] if bn else None, None, net.fuse_21, net.bn_fuse_21 if bn else None,
                          net.loc[1], net.conf[1], nc)]
    for k in range(4):
        blocks.append(SourceBlock(None, None, None, net.fuse_list1[k], net.bn_fuse_list1[k] if bn else None,
                                  net.loc[2 + k], net.conf[2 + k], nc))
    return blocks, (i43, i7)


def gssd_forward(net, x, detect_args=(0, 200, 0.01, 0.45), backbone=False):
    """Forward of the reference's SSD (ssd_multiphase_custom_group.py:217-400, ssd_type gssd) with every source chain
    — grouped conv / BN / ReLU / L2Norm / fuse 1x1 / BN / ReLU / loc+conf heads / permute / flatten / concat — run by
    the tcgen05 source blocks; the rest of the backbone stays the model's own torch modules.  Forward only
    (torch.no_grad): returns what `net(x)` returns — `(loc[B,P,4], conf[B,P,C], priors)` in the train phase,
    `Detect` output `[B,C,top_k,5]` in the test phase (ssd_multiphase_custom_group.py:382-396).

        net.forward = types.MethodType(gssd_forward, net)        # drop-in

    backbone=True (opt-in, eval mode) also runs every grouped backbone conv the kernel takes — conv3_2 .. conv5_3, with
    BatchNorm folded, ReLU fused and the pools on the PM layout — in bf16: 1.5x the model's fp32 torch forward at batch 32,
    at 1.0e-2 / 6.6e-3 relative error on loc / conf instead of 6.0e-3 / 3.7e-3 (tests/test_gpu_model.py).
    """
    import torch.nn.functional as F
    from ..functions import Detect
    _lib.require_cuda()
    if net.training and torch.is_grad_enabled() and not getattr(net, "_gssd_warned", False):
        import warnings
        warnings.warn("gssd_forward is forward-only: no autograd graph is recorded through the source blocks "
                      "(train-mode BatchNorm statistics are still updated)")
        net._gssd_warned = True
    cache = getattr(net, "_gssd_blocks", None)
    if cache is None:
        cache = build_source_blocks(net)
        net._gssd_blocks = cache
    blocks, (i43, i7) = cache
    bn = bool(net.batch_norm)
    with torch.no_grad():
        x = x.to(_lib.device_of(x))
        B = x.size(0)
        P = net.priors.size(0)
        loc = torch.empty((B, P, 4), dtype=torch.float32, device=x.device)
        conf = torch.empty((B, P, net.num_classes), dtype=torch.float32, device=x.device)
        off = 0
        if backbone and not net.training:
            # every grouped backbone conv the tcgen05 kernel takes (conv3_2 .. conv5_3) stays in PM/bf16 with BN folded
            run = getattr(net, "_gssd_backbone", None)
            if run is None:
                run = BackboneRun(net.vgg)
                net._gssd_backbone = run
            first = next((k for k in range(i43) if _conv_eligible(net.vgg[k])), i43)
            for k in range(first):                               # GSSD:254-259: the layers in front stay torch
                x = net.vgg[k](x)
            x = run(x, first, i43)
            x1, n = blocks[0](x, loc, conf, off)                 # conv4_3 .. heads of source 1 (GSSD:258-297, 375-377)
            off += n
            x = run(x1, i43 + (3 if bn else 2), i7)              # GSSD:300-301 up to the input of conv7
            x2, n = blocks[1](x, loc, conf, off)                 # conv7 .. heads of source 2 (GSSD:300-325)
            off += n
        else:
            for k in range(i43):                                 # GSSD:254-259, up to the input of conv4_3
                x = net.vgg[k](x)
            x1, n = blocks[0](x, loc, conf, off)                 # conv4_3 .. heads of source 1 (GSSD:258-297, 375-377)
            off += n
            x = x1.to_nchw()                                     # post-ReLU conv4_3 continues down the backbone
            for k in range(i43 + (3 if bn else 2), i7):          # GSSD:300-301 up to the input of conv7
                x = net.vgg[k](x)
            x2, n = blocks[1](x, loc, conf, off)                 # conv7 .. heads of source 2 (GSSD:300-325)
            off += n
        x = x2.to_nchw()
        si = 2
        for k, v in enumerate(net.extras):                       # GSSD:329-372
            x = v(x)
            if bn:
                if k % 2 == 1:
                    x = F.relu(x, inplace=True)
                is_source = k % 4 == 3
            else:
                x = F.relu(x, inplace=True)
                is_source = k % 2 == 1
            if is_source:
                _, n = blocks[si](x, loc, conf, off)
                off += n
                si += 1
        if off != P:
            raise RuntimeError("the sources produced %d priors, the model has %d" % (off, P))
        if net.phase == "test":                                  # GSSD:382-390
            return Detect.apply(net.num_classes, detect_args[0], detect_args[1], detect_args[2], detect_args[3],
                                loc, torch.softmax(conf, dim=-1), net.priors.to(x.device))
        return loc, conf, net.priors
