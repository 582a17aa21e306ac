"""SourceBlock — the per-source chain of the reference's SSD.forward
(models/ssd_multiphase_custom_group.py:258-297, 300-325, 329-372 and the heads at 375-380):

    grouped conv (groups=4) -> BN -> ReLU -> [L2Norm] -> fuse_X1 (1x1) -> bn_fuse_X1 -> ReLU
        -> loc.k / conf.k (3x3) -> permute(0,2,3,1) -> flatten -> concat

run as calls of libgssd_b200.so's tcgen05/TMEM implicit-GEMM convolution (`gssd_conv_igemm`).  The block does
not own parameters: it is built FROM the reference's own modules (`net.vgg[30]`, `net.vgg[31]`,
`net.L2Norm`, `net.fuse_11`, `net.bn_fuse_11`, `net.loc[0]`, `net.conf[0]` ...), so state-dict names and
checkpoints are the reference's.  Weights are re-packed to bf16 whenever a parameter changes.

With autograd enabled the chain is a `torch.autograd.Function` (`SourceBlock.forward_autograd`): the backward — data gradients on
the same implicit-GEMM kernel with re-packed filters, weight gradients on the split-K tcgen05 kernel `gssd_conv_wgrad`, BatchNorm
(batch statistics) / ReLU / L2Norm backward as row kernels on the PM layout — returns what autograd returns through the
reference's modules, so `gssd_forward` trains.
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
    anything else (conv6: dilation 6) round-trips through torch.  SURVEY §8f rank 1, evaluation mode; in training mode the same
    layers run under autograd as PMConvLayer nodes (run_layers(tc=...))."""

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

    def __init__(self, conv, groups, in_scale=None, extra=None, dev=None, pad_out=None):
        lib = _lib.require_cuda()
        ws = [conv.weight] + ([extra.weight] if extra is not None else [])
        bs = [conv.bias] + ([extra.bias] if extra is not None else [])
        w = torch.cat([_lib.f32(t, dev) for t in ws], 0)
        self.c_out_real = w.shape[0]
        if pad_out is not None and w.shape[0] % pad_out:                  # zero filters up to a multiple of `pad_out` output channels
            w = torch.cat([w, w.new_zeros((pad_out - w.shape[0] % pad_out,) + tuple(w.shape[1:]))], 0)
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
            if self.bias.shape[0] < self.c_out:
                self.bias = torch.cat([self.bias, self.bias.new_zeros((self.c_out - self.bias.shape[0],))], 0)
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
    Host-side half of the backward: packed with `gssd_conv_pack_weights`, it puts dgrad on the
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
            # heads: one launch that scatters straight into loc / conf while their 4A + A*C output channels fit one 64-column tile
            # (the reference's 2 classes: 24 / 36); wider heads (e.g. 21 VOC classes: 150) run as a plain convolution padded to a
            # multiple of 64 channels, followed by the permute / flatten of GSSD:376-380 in torch
            wide = self.loc.out_channels + self.conf.out_channels > 64
            h = _Conv(self.loc, 1, extra=self.conf, dev=dev, pad_out=64 if wide else None)
            h.wide = wide
            if not training:
                if g is not None and self.bn is not None:
                    g.fold_bn(self.bn)
                if self.bn_fuse is not None:
                    f.fold_bn(self.bn_fuse)
        self._packed, self._key = (g, f, h, training), key
        return self._packed

    @staticmethod
    def _bn_train(y, bn, stats, want_ss, out=None):
        """batch-statistics BatchNorm + ReLU on the raw conv output, in place or into `out` (the backward keeps the raw
        output); running stats as nn.BatchNorm2d"""
        lib = _lib.load()
        dev = y.data.device
        ss = torch.empty((y.rows,), dtype=torch.float32, device=dev) if want_ss else None
        mv = torch.empty((2 * y.c,), dtype=torch.float32, device=dev) if bn.track_running_stats else None
        gam = _lib.f32(bn.weight, dev) if bn.affine else None
        bet = _lib.f32(bn.bias, dev) if bn.affine else None
        with torch.cuda.device(dev):
            if out is None:
                _lib.check(lib.gssd_bn_act_pm(y.data.data_ptr(), y.n, y.c, y.h, y.w, stats.data_ptr(), _lib.ptr(gam), _lib.ptr(bet),
                                              float(bn.eps), 1, _lib.ptr(ss), _lib.ptr(mv), _lib.stream()), "gssd_bn_act_pm")
            else:
                _lib.check(lib.gssd_bn_act_pm_to(y.data.data_ptr(), out.data.data_ptr(), y.n, y.c, y.h, y.w, stats.data_ptr(),
                                                 _lib.ptr(gam), _lib.ptr(bet), float(bn.eps), 1, _lib.ptr(ss), _lib.ptr(mv),
                                                 _lib.stream()), "gssd_bn_act_pm_to")
        if mv is not None:
            with torch.no_grad():
                bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - mom).add_(mv[:y.c].to(bn.running_mean.device), alpha=mom)
                bn.running_var.mul_(1 - mom).add_(mv[y.c:].to(bn.running_var.device), alpha=mom)
        return ss

    def param_list(self):
        """(name, tensor or None) of every parameter of the chain, in the order _SourceChainFn takes them"""
        def wb(m, n):
            return [(n + "_w", m.weight if m is not None else None), (n + "_b", m.bias if m is not None else None)]
        return (wb(self.gconv, "gconv") + wb(self.bn if (self.bn is not None and self.bn.affine) else None, "bn") +
                [("l2norm_w", self.l2norm.weight if self.l2norm is not None else None)] + wb(self.fuse, "fuse") +
                wb(self.bn_fuse if (self.bn_fuse is not None and self.bn_fuse.affine) else None, "bn_fuse") +
                wb(self.loc, "loc") + wb(self.conf, "conf"))

    def _pack_dgrad(self, dev):
        """the three data-gradient convolutions (heads, fuse, grouped conv), re-packed whenever a parameter changes"""
        mods = (self.gconv, self.l2norm, self.fuse, self.loc, self.conf)
        key = (_versions(*mods), str(dev))
        if getattr(self, "_dg_key", None) != key:
            with torch.no_grad(), torch.cuda.device(dev):
                wh = torch.cat([_lib.f32(self.loc.weight, dev), _lib.f32(self.conf.weight, dev)], 0)
                d = dict(h=_DgradConv(wh, 1, dev, pad_in=64),
                         f=_DgradConv(self.fuse.weight, 1, dev, in_scale=self.l2norm.weight if self.l2norm is not None else None))
                if self.gconv is not None:
                    d["g"] = _DgradConv(self.gconv.weight, self.gconv.groups, dev)
            self._dg, self._dg_key = d, key
        return self._dg

    def forward_autograd(self, x):
        """the chain as one autograd node: x NCHW fp32 -> (loc_k[B, n_k, 4], conf_k[B, n_k, C], x_out NCHW or None).  BatchNorm
        in training mode (batch statistics), in eval mode (running statistics, folded into the conv epilogue) or absent."""
        prm = [t for _, t in self.param_list()]
        out = _SourceChainFn.apply(self, x, *prm)
        return (out[0], out[1], out[2] if len(out) > 2 else None)

    def _heads(self, src, h, loc_out, conf_out, prior_off, P):
        """loc.k / conf.k on `src` (PM) -> this source's slice of loc_out[B,P,4] / conf_out[B,P,C] (GSSD:375-380)"""
        if not h.wide:
            conv_igemm(src, h, relu=False, shift=h.bias, head=(loc_out, conf_out, self.n_anchor, self.num_classes, prior_off, P))
            return
        y = conv_igemm(src, h, relu=False, shift=h.bias).to_nchw()[:, :h.c_out_real].permute(0, 2, 3, 1)   # [N, H, W, 4A + A*C]
        n_k = src.h * src.w * self.n_anchor
        la = 4 * self.n_anchor
        loc_out[:, prior_off:prior_off + n_k] = y[..., :la].reshape(src.n, n_k, 4)
        conf_out[:, prior_off:prior_off + n_k] = y[..., la:].reshape(src.n, n_k, self.num_classes)

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
            self._heads(src, h, loc_out, conf_out, int(prior_off), P)
        return x1, x.h * x.w * self.n_anchor

    __call__ = forward



# ---- the chain under autograd -------------------------------------------------------------------------------------
def _wgrad(dy, x, c_out, groups, taps):
    """weight gradient of a stride-1 'same' convolution: dy PM [rows, >= c_out], x PM [rows, c_in] -> fp32 [c_out, c_in/groups, k, k]"""
    lib = _lib.load()
    dev = x.data.device
    cg = x.c // groups
    nbytes = lib.gssd_conv_wgrad_bytes(x.c, c_out, groups, taps)
    dw = torch.empty((nbytes // 4,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.gssd_conv_wgrad(dy.data.data_ptr(), x.data.data_ptr(), x.n, x.h, x.w, x.c, c_out, dy.c, groups, taps,
                                       dw.data_ptr(), _lib.stream()), "gssd_conv_wgrad")
    k = 3 if taps == 9 else 1
    dw = dw.view(taps, -1, cg)[:, :c_out]                                # [taps, c_out, cg]
    return dw.permute(1, 2, 0).reshape(c_out, cg, k, k)


class _DgradConv(object):
    """the convolution that computes the data gradient of `weight` (see dgrad_weight), packed for gssd_conv_igemm; the input
    channels (= output channels of the forward conv) are zero-padded to `pad_in` for the heads"""

    def __init__(self, weight, groups, dev, in_scale=None, pad_in=None):
        import torch.nn as nn
        w = _lib.f32(weight, dev)
        if in_scale is not None:                                         # L2Norm.weight folded into the forward filter
            w = w * _lib.f32(in_scale, dev).view(1, -1, 1, 1)
        wd = dgrad_weight(w, groups)                                     # [c_in, c_out/groups, k, k]
        if pad_in is not None and wd.shape[1] < pad_in:
            wd = torch.cat([wd, wd.new_zeros((wd.shape[0], pad_in - wd.shape[1]) + tuple(wd.shape[2:]))], 1)
        k = wd.shape[2]
        conv = nn.Conv2d(wd.shape[1] * groups, wd.shape[0], k, padding=k // 2, groups=groups, bias=False).to(dev)
        with torch.no_grad():
            conv.weight.copy_(wd)
        self.cv = _Conv(conv, groups, dev=dev)


def _bn_relu_bwd(dy, y, yraw, add, stats, gamma, bn_eps, ss_l2, l2_eps, ss_out, l2_eps_out, ebn=None):
    """gssd_bn_relu_bwd_pm -> (dx PM, sums[3c] = (dbeta, dgamma, dbias))"""
    lib = _lib.load()
    dev = dy.data.device
    out = PM.empty(dy.n, dy.c, dy.h, dy.w, dev)
    sums = torch.empty((3 * dy.c,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.gssd_bn_relu_bwd_pm(dy.data.data_ptr(), y.data.data_ptr(), _lib.ptr(yraw.data if yraw is not None else None),
                                           _lib.ptr(add.data if add is not None else None), dy.n, dy.c, dy.h, dy.w,
                                           _lib.ptr(stats), _lib.ptr(gamma), float(bn_eps), 1,
                                           _lib.ptr(ebn[0] if ebn else None), _lib.ptr(ebn[1] if ebn else None), _lib.ptr(ss_l2), float(l2_eps),
                                           _lib.ptr(ss_out), float(l2_eps_out), out.data.data_ptr(), sums.data_ptr(), _lib.stream()),
                   "gssd_bn_relu_bwd_pm")
    return out, sums


class _SourceChainFn(torch.autograd.Function):
    """One source chain under autograd.  forward(blk, x, *params) -> (loc_k[B, n_k, 4], conf_k[B, n_k, C], x_out NCHW or None);
    `params` = the tensors of SourceBlock.param_list() (only there so that autograd routes their gradients)."""

    @staticmethod
    def forward(ctx, blk, x, *params):
        dev = _lib.device_of(x)
        with torch.cuda.device(dev):
            x0 = x if isinstance(x, PM) else PM.from_nchw(x)
            g, f, h, training = blk._pack(dev)
            if h.wide:
                raise NotImplementedError("SourceBlock under autograd: heads wider than 64 output channels (%d anchors x (4 + %d classes)); "
                                          "the forward supports them, the backward is written for the one-tile heads of GSSD"
                                          % (blk.n_anchor, blk.num_classes))
            want_l2 = blk.l2norm is not None
            eps = blk.l2norm.eps if want_l2 else 0.0
            sv = dict(x0=x0, training=training, l2=want_l2, eps=eps)
            ss = torch.empty((x0.rows,), dtype=torch.float32, device=dev) if want_l2 else None
            if g is not None:
                if training and blk.bn is not None:
                    stats1 = torch.zeros((2 * g.c_out,), dtype=torch.float32, device=dev)
                    y1raw = conv_igemm(x0, g, relu=False, shift=g.bias, chan_sum=stats1)
                    y1 = PM.empty(y1raw.n, y1raw.c, y1raw.h, y1raw.w, dev)
                    ss = blk._bn_train(y1raw, blk.bn, stats1, want_l2, out=y1)
                    sv.update(y1raw=y1raw, stats1=stats1)
                else:
                    y1 = conv_igemm(x0, g, relu=True, scale=g.scale, shift=g.shift, row_ss_out=ss)
            else:
                if want_l2:
                    raise NotImplementedError("L2Norm without the grouped conv in front")
                y1 = x0
            if training and blk.bn_fuse is not None:
                stats2 = torch.zeros((2 * f.c_out,), dtype=torch.float32, device=dev)
                y2raw = conv_igemm(y1, f, relu=False, shift=f.bias, chan_sum=stats2, row_ss_in=ss, l2_eps=eps)
                z2 = PM.empty(y2raw.n, y2raw.c, y2raw.h, y2raw.w, dev)
                blk._bn_train(y2raw, blk.bn_fuse, stats2, False, out=z2)
                sv.update(y2raw=y2raw, stats2=stats2)
            else:
                z2 = conv_igemm(y1, f, relu=True, scale=f.scale, shift=f.shift, row_ss_in=ss, l2_eps=eps)
            n_k = x0.h * x0.w * blk.n_anchor
            loc = torch.empty((x0.n, n_k, 4), dtype=torch.float32, device=dev)
            conf = torch.empty((x0.n, n_k, blk.num_classes), dtype=torch.float32, device=dev)
            conv_igemm(z2, h, relu=False, shift=h.bias, head=(loc, conf, blk.n_anchor, blk.num_classes, 0, n_k))
            sv.update(y1=y1, z2=z2, ss=ss, n_k=n_k)
            x_out = y1.to_nchw() if g is not None else None
        ctx.blk, ctx.sv = blk, sv
        ctx.x_is_pm = isinstance(x, PM)
        if x_out is None:
            return loc, conf
        return loc, conf, x_out

    @staticmethod
    def backward(ctx, d_loc, d_conf, d_xout=None):
        blk, sv = ctx.blk, ctx.sv
        lib = _lib.load()
        x0, y1, z2, ss = sv["x0"], sv["y1"], sv["z2"], sv["ss"]
        dev = x0.data.device
        A, NC = blk.n_anchor, blk.num_classes
        c_head = A * (4 + NC)
        training = sv["training"]
        with torch.no_grad(), torch.cuda.device(dev):
            g, f, h, _ = blk._pack(dev)
            dg = blk._pack_dgrad(dev)
            grads = {}
            # ---- heads -----------------------------------------------------------------------------------------------------
            if d_loc is None:
                d_loc = torch.zeros((x0.n, sv["n_k"], 4), dtype=torch.float32, device=dev)
            if d_conf is None:
                d_conf = torch.zeros((x0.n, sv["n_k"], NC), dtype=torch.float32, device=dev)
            d_loc, d_conf = _lib.f32(d_loc, dev), _lib.f32(d_conf, dev)
            dH = PM.empty(x0.n, 64, x0.h, x0.w, dev)
            bias_h = torch.empty((c_head,), dtype=torch.float32, device=dev)
            _lib.check(lib.gssd_head_grad_pm(d_loc.data_ptr(), d_conf.data_ptr(), x0.n, x0.h, x0.w, sv["n_k"], 0, A, NC, 64,
                                             dH.data.data_ptr(), bias_h.data_ptr(), _lib.stream()), "gssd_head_grad_pm")
            dWh = _wgrad(dH, z2, c_head, 1, 9)
            grads["loc_w"], grads["conf_w"] = dWh[:4 * A], dWh[4 * A:]
            grads["loc_b"], grads["conf_b"] = bias_h[:4 * A], bias_h[4 * A:]
            dz2 = conv_igemm(dH, dg["h"].cv, relu=False)
            # ---- bn_fuse + ReLU (+ the row factor of the deferred L2Norm of the fuse input) ------------------------------
            bnf = blk.bn_fuse
            if training and bnf is not None:
                t2, s2 = _bn_relu_bwd(dz2, z2, sv["y2raw"], None, sv["stats2"], _lib.f32(bnf.weight, dev) if bnf.affine else None,
                                      bnf.eps, None, 0.0, ss, sv["eps"])
                grads["bn_fuse_b"], grads["bn_fuse_w"] = s2[:f.c_out], s2[f.c_out:2 * f.c_out]
            else:
                ebn = (_lib.f32(bnf.weight, dev), _lib.f32(bnf.bias, dev)) if (bnf is not None and bnf.affine) else None
                t2, s2 = _bn_relu_bwd(dz2, z2, None, None, None, f.scale, 0.0, None, 0.0, ss, sv["eps"], ebn=ebn)
                if ebn is not None:
                    grads["bn_fuse_b"], grads["bn_fuse_w"] = s2[:f.c_out], s2[f.c_out:2 * f.c_out]
            grads["fuse_b"] = s2[2 * f.c_out:]
            dWf = _wgrad(t2, y1, f.c_out, 1, 1)                          # w.r.t. the filter as packed (L2Norm.weight folded in)
            if sv["l2"]:
                l2w = _lib.f32(blk.l2norm.weight, dev)
                grads["l2norm_w"] = (dWf * _lib.f32(blk.fuse.weight, dev)).sum(dim=(0, 2, 3))
                grads["fuse_w"] = dWf * l2w.view(1, -1, 1, 1)
            else:
                grads["fuse_w"] = dWf
            a1 = conv_igemm(t2, dg["f"].cv, relu=False)
            # ---- bn + ReLU of the grouped conv (+ the L2Norm backward, + what flows back from further down the backbone) ------
            if g is None:
                dx = a1
            else:
                add = PM.from_nchw(d_xout) if d_xout is not None else None
                bn = blk.bn
                if training and bn is not None:
                    d1, s1 = _bn_relu_bwd(a1, y1, sv["y1raw"], add, sv["stats1"], _lib.f32(bn.weight, dev) if bn.affine else None,
                                          bn.eps, ss if sv["l2"] else None, sv["eps"], None, 0.0)
                    grads["bn_b"], grads["bn_w"] = s1[:g.c_out], s1[g.c_out:2 * g.c_out]
                else:
                    ebn = (_lib.f32(bn.weight, dev), _lib.f32(bn.bias, dev)) if (bn is not None and bn.affine) else None
                    d1, s1 = _bn_relu_bwd(a1, y1, None, add, None, g.scale, 0.0, ss if sv["l2"] else None, sv["eps"], None, 0.0, ebn=ebn)
                    if ebn is not None:
                        grads["bn_b"], grads["bn_w"] = s1[:g.c_out], s1[g.c_out:2 * g.c_out]
                grads["gconv_b"] = s1[2 * g.c_out:]
                grads["gconv_w"] = _wgrad(d1, x0, g.c_out, g.groups, g.taps)
                dx = conv_igemm(d1, dg["g"].cv, relu=False)
            gx = dx if ctx.x_is_pm else dx.to_nchw()
            if getattr(blk, "_debug_backward", None) is not None:        # development: the intermediate gradients, as NCHW fp32
                blk._debug_backward.update(dH=dH.to_nchw(), dz2=dz2.to_nchw(), t2=t2.to_nchw(), a1=a1.to_nchw(), z2=z2.to_nchw(),
                                           y1=y1.to_nchw(), ss=ss)
                if g is not None:
                    blk._debug_backward.update(d1=d1.to_nchw())
        out = [None, gx if ctx.needs_input_grad[1] else None]
        for name, prm in blk.param_list():
            gr = grads.get(name)
            if gr is not None and prm is not None:
                gr = gr.reshape(prm.shape).to(prm.dtype)
            out.append(gr)
        return tuple(out)


# ---- the grouped backbone convolutions under autograd (second half) -------------------------------------------------
# conv3_2 .. conv5_3 of the reference's vgg list (ssd_multiphase_custom_group.py:434-460) are [Conv2d(groups=4), BatchNorm2d, ReLU]
# triples with 64 or 128 channels per group: each runs as one autograd node on the kernels of the source block — forward
# gssd_conv_igemm + gssd_bn_act_pm_to, backward gssd_bn_relu_bwd_pm + gssd_conv_wgrad + gssd_conv_igemm on the rotated filter — with
# bf16 PM tensors (and bf16 PM gradients) flowing between consecutive triples.
class _NchwToPMFn(torch.autograd.Function):
    """NCHW fp32 -> the bf16 [rows, c] tensor of a PM; the gradient comes back as such a tensor and leaves as NCHW fp32"""

    @staticmethod
    def forward(ctx, x):
        pm = PM.from_nchw(x)
        ctx.geom = (pm.n, pm.c, pm.h, pm.w)
        return pm.data

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d):
        return PM(d.contiguous(), *ctx.geom).to_nchw()


class _PMToNchwFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, n, c, h, w):
        return PM(data, n, c, h, w).to_nchw()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        return PM.from_nchw(dy).data, None, None, None, None


class _Filter(object):
    """what _Conv reads of an nn.Conv2d, for filters that are not a module's parameter (the rotated filter of a data gradient)"""

    def __init__(self, weight, k):
        self.weight, self.bias = weight, None
        self.kernel_size, self.stride, self.dilation, self.padding = (k, k), (1, 1), (1, 1), (k // 2, k // 2)


def _wgrad_any(dy, x, c_out, groups, taps):
    """_wgrad for every group width the forward kernel takes.  gssd_conv_wgrad's tiles are 128 channels wide; with 64 channels per
    group, neighbouring groups are handed to it as ONE group of 128 and the cross-group blocks it computes on the way (the gradient
    of filter entries a grouped convolution does not have) are dropped."""
    cg, ng = x.c // groups, c_out // groups
    if cg % 128 == 0 and (groups == 1 or ng % 128 == 0):
        return _wgrad(dy, x, c_out, groups, taps)
    m = 128 // cg if (cg > 0 and 128 % cg == 0) else 0
    if m < 2 or groups % m or (m * ng) % 128:
        raise NotImplementedError("weight gradient for %d -> %d channels in %d groups" % (x.c, c_out, groups))
    full = _wgrad(dy, x, c_out, groups // m, taps)                        # [c_out, m*cg, k, k]
    k = full.shape[-1]
    t = full.view(groups // m, m, ng, m, cg, k, k)
    return torch.stack([t[:, j, :, j] for j in range(m)], 1).reshape(c_out, cg, k, k)


class PMConvLayer(object):
    """one (Conv2d, BatchNorm2d in training mode, ReLU) triple of the backbone on PM tensors"""
    calls = 0                                                             # forward calls so far (tests / bench: the path was taken)
    debug_outputs = None                                                  # tests: a list that receives every forward's output (NCHW fp32)

    def __init__(self, conv, bn):
        self.conv, self.bn = conv, bn
        self._fwd = self._dg = None

    @staticmethod
    def takes(conv, bn, relu, x=None, width=None):
        """x: the NCHW input of the triple (checked when given); width: the feature-map width when only that is known"""
        import torch.nn as nn
        if not (_conv_eligible(conv) and isinstance(bn, nn.BatchNorm2d) and isinstance(relu, nn.ReLU)):
            return False
        if not bn.training or bn.num_features != conv.out_channels or conv.padding_mode != "zeros":
            return False
        cg, ng, g = conv.in_channels // conv.groups, conv.out_channels // conv.groups, conv.groups
        wgrad_ok = (cg % 128 == 0 and (g == 1 or ng % 128 == 0)) or (cg == 64 and g % 2 == 0 and (2 * ng) % 128 == 0)
        if not wgrad_ok or conv.out_channels % 256 or conv.out_channels > 1024:        # gssd_bn_relu_bwd_pm: whole 256-channel rows
            return False
        # widest feature map whose 3x3 slab ring fits twice beside the weight stages of gssd_conv_igemm (forward and data gradient;
        # its host-side stage arithmetic gives 61 pixels per row with 128-column weight tiles, 157 with 64-column tiles)
        w_max = 60 if (ng % 128 == 0 or cg % 128 == 0) else 144
        if width is not None and width > w_max:
            return False
        return x is None or (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
                             and x.shape[1] == conv.in_channels and x.shape[3] <= w_max)

    def fwd(self, dev):
        key = (_versions(self.conv), str(dev))
        if self._fwd is None or self._fwd[0] != key:
            with torch.no_grad(), torch.cuda.device(dev):
                self._fwd = (key, _Conv(self.conv, self.conv.groups, dev=dev))
        return self._fwd[1]

    def dgrad(self, dev):
        key = (_versions(self.conv), str(dev))
        if self._dg is None or self._dg[0] != key:
            with torch.no_grad(), torch.cuda.device(dev):
                g = self.conv.groups
                wd = dgrad_weight(_lib.f32(self.conv.weight, dev), g)
                self._dg = (key, _Conv(_Filter(wd, wd.shape[2]), g, dev=dev))
        return self._dg[1]

    def __call__(self, xdata, n, h, w):
        """xdata: bf16 [n*(h+2)*(w+2), c_in] -> bf16 [rows, c_out] (the ReLU output), recorded by autograd"""
        conv, bn = self.conv, self.bn
        gamma, beta = (bn.weight, bn.bias) if bn.affine else (None, None)
        return _PMConvBnReluFn.apply(xdata, self, n, h, w, conv.weight, conv.bias, gamma, beta)


class _PMConvBnReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xdata, layer, n, h, w, weight, bias, gamma, beta):
        dev = xdata.device
        conv, bn = layer.conv, layer.bn
        with torch.cuda.device(dev):
            x0 = PM(xdata, n, conv.in_channels, h, w)
            cv = layer.fwd(dev)
            stats = torch.zeros((2 * cv.c_out,), dtype=torch.float32, device=dev)
            yraw = conv_igemm(x0, cv, relu=False, shift=cv.bias, chan_sum=stats)
            y = PM.empty(n, cv.c_out, h, w, dev)
            SourceBlock._bn_train(yraw, bn, stats, False, out=y)
        PMConvLayer.calls += 1
        if PMConvLayer.debug_outputs is not None:
            PMConvLayer.debug_outputs.append(y.to_nchw())
        ctx.layer, ctx.geom = layer, (n, h, w)
        ctx.save_for_backward(xdata, yraw.data, y.data, stats)
        return y.data

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        xdata, yraw_d, y_d, stats = ctx.saved_tensors
        layer = ctx.layer
        n, h, w = ctx.geom
        conv, bn = layer.conv, layer.bn
        dev = xdata.device
        c_in, c_out = conv.in_channels, conv.out_channels
        with torch.cuda.device(dev):
            x0, yraw, y = PM(xdata, n, c_in, h, w), PM(yraw_d, n, c_out, h, w), PM(y_d, n, c_out, h, w)
            d = PM(dy.to(torch.bfloat16).contiguous(), n, c_out, h, w)
            gam = _lib.f32(bn.weight, dev) if bn.affine else None
            d1, s = _bn_relu_bwd(d, y, yraw, None, stats, gam, bn.eps, None, 0.0, None, 0.0)
            dw = _wgrad_any(d1, x0, c_out, conv.groups, conv.kernel_size[0] * conv.kernel_size[1])
            dx = conv_igemm(d1, layer.dgrad(dev), relu=False).data if ctx.needs_input_grad[0] else None
        # the bias of a convolution in front of a training-mode BatchNorm has no gradient (the batch mean removes any per-channel
        # constant); s[2c:] = sum of dx holds only the rounding noise of that zero
        d_bias = torch.zeros_like(conv.bias) if conv.bias is not None else None
        d_gamma = s[c_out:2 * c_out].to(bn.weight.dtype) if bn.affine else None
        d_beta = s[:c_out].to(bn.bias.dtype) if bn.affine else None
        return dx, None, None, None, None, dw.to(conv.weight.dtype), d_bias, d_gamma, d_beta


def pm_layers_at(modules, k, stop, x):
    """the run of consecutive (Conv2d, BatchNorm2d, ReLU) triples PMConvLayer takes, starting at modules[k]: -> list of (conv, bn)"""
    run = []
    c = None
    if not (isinstance(x, torch.Tensor) and x.dim() == 4):
        return run
    while k + 3 <= stop:
        conv, bn, relu = modules[k], modules[k + 1], modules[k + 2]
        if not PMConvLayer.takes(conv, bn, relu, x if not run else None, x.shape[3]) or (run and conv.in_channels != c):
            break
        run.append((conv, bn))
        c = conv.out_channels
        k += 3
    return run


def run_pm_layers(pairs, x, cache):
    """x (NCHW fp32) through a run of triples (see pm_layers_at); `cache`: dict that keeps the PMConvLayer of every conv"""
    n, _, h, w = x.shape
    data = _NchwToPMFn.apply(x)
    c = x.shape[1]
    for conv, bn in pairs:
        layer = cache.get(id(conv))
        if layer is None or layer.conv is not conv or layer.bn is not bn:
            layer = PMConvLayer(conv, bn)
            cache[id(conv)] = layer
        data = layer(data, n, h, w)
        c = conv.out_channels
    return _PMToNchwFn.apply(data, n, c, h, w)


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
    the tcgen05 source blocks; the rest of the backbone stays the model's own torch modules.  With autograd enabled the
    source chains are autograd nodes (their backward runs on the same kernels), so `loss.backward()` reaches every
    parameter exactly as through the reference forward; returns what `net(x)` returns — `(loc[B,P,4], conf[B,P,C], priors)` in the train phase,
    `Detect` output `[B,C,top_k,5]` in the test phase (ssd_multiphase_custom_group.py:382-396).

        net.forward = types.MethodType(gssd_forward, net)        # drop-in

    backbone=True (opt-in) also runs every grouped backbone conv the kernel takes — conv3_2 .. conv5_3 — in bf16.  Eval mode without
    autograd: BatchNorm folded, ReLU fused and the pools on the PM layout (BackboneRun): 1.5x the model's fp32 torch forward at
    batch 32, at 1.0e-2 / 6.6e-3 relative error on loc / conf instead of 6.0e-3 / 3.7e-3 (tests/test_gpu_model.py).  Training mode:
    every (Conv2d, BatchNorm2d, ReLU) triple of that stretch is one autograd node on the same kernels as the source blocks
    (PMConvLayer: batch statistics, data gradient on the forward kernel, weight gradient on gssd_conv_wgrad), the pools stay torch's.
    """
    import contextlib
    import os
    import torch.nn.functional as F
    from ..functions import Detect
    from .bn_relu import bn_relu, run_layers, takes as bn_takes, to_channels_last
    _lib.require_cuda()
    fuse_bn = os.environ.get("GSSD_FUSED_BN", "1") != "0"    # development: A/B against torch's BatchNorm2d + ReLU modules
    # opt-in (GSSD_CHANNELS_LAST=1): the torch convolutions between the source blocks run channels-last (cuDNN's NHWC kernels, no
    # layout conversion passes), with the fused BatchNorm / ReLU / pool kernels in their channels-last form.  Measured at batch 32:
    # 33.4 ms against 33.9 ms per training step - what cuDNN saves on conversions it loses on its NHWC grouped convolutions - so it
    # is not the default
    chl = fuse_bn and net.training and os.environ.get("GSSD_CHANNELS_LAST", "0") == "1"

    def _plain(mods, xx, a, b):
        for kk in range(a, b):
            xx = mods[kk](xx)
        return xx
    cache = getattr(net, "_gssd_blocks", None)
    if cache is None:
        cache = build_source_blocks(net)
        net._gssd_blocks = cache
    blocks, (i43, i7) = cache
    bn = bool(net.batch_norm)
    # under autograd every source chain is one autograd node (SourceBlock.forward_autograd) and the layers in between are the
    # model's own modules, recorded by torch as usual; without it everything runs under no_grad
    use_ag = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in net.parameters()))
    with (contextlib.nullcontext() if use_ag else torch.no_grad()):
        x = x.to(_lib.device_of(x))
        B = x.size(0)
        P = net.priors.size(0)
        off = 0
        if use_ag:
            locs, confs = [], []

            def run_block(blk, xin):
                nonlocal off
                l, c, xo = blk.forward_autograd(xin)
                locs.append(l); confs.append(c)
                off += l.shape[1]
                return xo
            loc = conf = None
        else:
            loc = torch.empty((B, P, 4), dtype=torch.float32, device=x.device)
            conf = torch.empty((B, P, net.num_classes), dtype=torch.float32, device=x.device)

            def run_block(blk, xin):
                nonlocal off
                xo, n = blk(xin, loc, conf, off)
                off += n
                return xo
        if backbone and not net.training and not use_ag:
            # every grouped backbone conv the tcgen05 kernel takes (conv3_2 .. conv5_3) stays in PM/bf16 with BN folded
            run = getattr(net, "_gssd_backbone", None)
            if run is None:
                run = BackboneRun(net.vgg)
                net._gssd_backbone = run
            first = next((k for k in range(i43) if _conv_eligible(net.vgg[k])), i43)
            for k in range(first):                               # GSSD:254-259: the layers in front stay torch
                x = net.vgg[k](x)
            x = run(x, first, i43)
            x1 = run_block(blocks[0], x)                         # conv4_3 .. heads of source 1 (GSSD:258-297, 375-377)
            x = run(x1, i43 + (3 if bn else 2), i7)              # GSSD:300-301 up to the input of conv7
            x2 = run_block(blocks[1], x)                         # conv7 .. heads of source 2 (GSSD:300-325)
        else:
            # the backbone layers in between stay the model's torch convolutions / pools; in training mode every (BatchNorm2d, ReLU)
            # pair behind them runs as one fused node (bn_relu.py: the largest cost of the reference's training step)
            if chl:
                if not getattr(net, "_gssd_channels_last", False):
                    to_channels_last(net.vgg)
                    net._gssd_channels_last = True
                x = x.contiguous(memory_format=torch.channels_last)
            # backbone=True in training mode: conv3_2 .. conv5_3 on the tcgen05 kernels too, forward and backward (PMConvLayer)
            tc = None
            if backbone and net.training and fuse_bn:
                tc = getattr(net, "_gssd_pm_layers", None)
                if tc is None:
                    tc = net._gssd_pm_layers = {}
            x = run_layers(net.vgg, x, 0, i43, tc=tc) if fuse_bn else _plain(net.vgg, x, 0, i43)      # GSSD:254-259, up to the input of conv4_3
            x1 = run_block(blocks[0], x)                         # conv4_3 .. heads of source 1 (GSSD:258-297, 375-377)
            x = x1 if use_ag else x1.to_nchw()                   # post-ReLU conv4_3 continues down the backbone
            if chl:
                x = x.contiguous(memory_format=torch.channels_last)
            k0 = i43 + (3 if bn else 2)                          # GSSD:300-301 up to the input of conv7
            x = run_layers(net.vgg, x, k0, i7, tc=tc) if fuse_bn else _plain(net.vgg, x, k0, i7)
            x2 = run_block(blocks[1], x)                         # conv7 .. heads of source 2 (GSSD:300-325)
        x = x2 if use_ag else x2.to_nchw()
        si = 2
        for k, v in enumerate(net.extras):                       # GSSD:329-372
            if bn and k % 2 == 1 and fuse_bn and bn_takes(x, v):
                x = bn_relu(x, v, relu=True)
                is_source = k % 4 == 3
            elif bn:
                x = v(x)
                if k % 2 == 1:
                    x = F.relu(x, inplace=True)
                is_source = k % 4 == 3
            else:
                x = F.relu(v(x), inplace=True)
                is_source = k % 2 == 1
            if is_source:
                run_block(blocks[si], x)
                si += 1
        if off != P:
            raise RuntimeError("the sources produced %d priors, the model has %d" % (off, P))
        if use_ag:
            loc, conf = torch.cat(locs, 1), torch.cat(confs, 1)
        if net.phase == "test":                                  # GSSD:382-390: softmax(conf) evaluated inside Detect's threshold pass
            return Detect.apply_logits(net.num_classes, detect_args[0], detect_args[1], detect_args[2], detect_args[3],
                                       loc, conf, net.priors.to(x.device))
        return loc, conf, net.priors
