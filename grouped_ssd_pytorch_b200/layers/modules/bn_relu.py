"""Training-mode `nn.BatchNorm2d` + `F.relu` of the backbone layers that stay NCHW fp32 torch convolutions
(models/ssd_multiphase_custom_group.py:434-460, 254-259, 300-301) as one autograd node on `gssd_bn_relu_nchw_fwd / _bwd`:
two streaming kernels forward, two backward, the ReLU mask recomputed from the input — against cuDNN's batch-norm kernels plus the
separate ReLU / threshold_backward passes torch runs (the largest cost of the reference's batch-32 training step on a B200).

`bn_relu(x, bn, relu=True)` uses the module's own parameters and updates its running statistics / `num_batches_tracked` exactly as
`bn(x)` would; evaluation mode, CPU tensors, non-fp32 inputs and `momentum=None` modules are not taken (the caller keeps torch's
modules for those).  `run_layers(modules, x)` walks a slice of an `nn.ModuleList` / `nn.Sequential` and fuses every
(BatchNorm2d, ReLU) pair it meets."""
import torch
import torch.nn as nn

from ... import _lib


def _is_channels_last(x):
    """a genuinely channels-last 4-D tensor (not one that is contiguous in both senses, e.g. 1x1 maps or one channel)"""
    return x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)


def _nhwc_ok(c):
    return c % 4 == 0 and c <= 1024 and 256 % (c // 4) == 0


class _BnReluTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu, conv_bias=None):
        lib = _lib.require_cuda()
        dev = x.device
        cl = _is_channels_last(x) and _nhwc_ok(x.shape[1])
        xc = x if cl else x.contiguous()
        N, C = xc.shape[0], xc.shape[1]
        HW = xc.numel() // (N * C)
        y = torch.empty_like(xc)                                  # keeps the memory format
        save = torch.empty((2 * C,), dtype=torch.float32, device=dev)
        ws = torch.empty((2 * C,), dtype=torch.float64, device=dev)
        w, b = _lib.f32(weight, dev), _lib.f32(bias, dev)
        shift = None if conv_bias is None else _lib.f32(conv_bias, dev)
        with torch.cuda.device(dev):
            if cl:
                _lib.check(lib.gssd_bn_relu_nhwc_fwd(xc.data_ptr(), w.data_ptr(), b.data_ptr(), N * HW, C, float(eps), 1 if relu else 0,
                                                     y.data_ptr(), save.data_ptr(), _lib.ptr(running_mean), _lib.ptr(running_var),
                                                     float(momentum), _lib.ptr(shift), ws.data_ptr(), _lib.stream()), "gssd_bn_relu_nhwc_fwd")
            else:
                _lib.check(lib.gssd_bn_relu_nchw_fwd(xc.data_ptr(), w.data_ptr(), b.data_ptr(), N, C, HW, float(eps), 1 if relu else 0,
                                                     y.data_ptr(), save.data_ptr(), _lib.ptr(running_mean), _lib.ptr(running_var),
                                                     float(momentum), _lib.ptr(shift), ws.data_ptr(), _lib.stream()), "gssd_bn_relu_nchw_fwd")
        ctx.save_for_backward(xc, w, b, save)
        ctx.relu, ctx.cl = relu, cl
        ctx.conv_bias = None if conv_bias is None else (tuple(conv_bias.shape), conv_bias.dtype)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        lib = _lib.require_cuda()
        xc, w, b, save = ctx.saved_tensors
        dev = xc.device
        N, C = xc.shape[0], xc.shape[1]
        HW = xc.numel() // (N * C)
        dyc = dy.detach().to(device=dev, dtype=torch.float32)
        dyc = dyc.contiguous(memory_format=torch.channels_last) if ctx.cl else dyc.contiguous()
        dx = torch.empty_like(xc)
        dg, db = torch.empty_like(w), torch.empty_like(b)
        ws = torch.empty((2 * C,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            if ctx.cl:
                _lib.check(lib.gssd_bn_relu_nhwc_bwd(xc.data_ptr(), dyc.data_ptr(), w.data_ptr(), b.data_ptr(), save.data_ptr(), N * HW, C,
                                                     1 if ctx.relu else 0, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                                     _lib.stream()), "gssd_bn_relu_nhwc_bwd")
            else:
                _lib.check(lib.gssd_bn_relu_nchw_bwd(xc.data_ptr(), dyc.data_ptr(), w.data_ptr(), b.data_ptr(), save.data_ptr(), N, C, HW,
                                                     1 if ctx.relu else 0, dx.data_ptr(), dg.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                                     _lib.stream()), "gssd_bn_relu_nchw_bwd")
        # a left-out convolution bias has no gradient (training-mode BN removes any per-channel constant; torch's own backward leaves
        # rounding noise of the order 1e-6 there)
        d_cb = None if ctx.conv_bias is None else torch.zeros(ctx.conv_bias[0], dtype=ctx.conv_bias[1], device=dev)
        return dx, dg, db, None, None, None, None, None, d_cb


def takes(x, bn):
    """whether `bn_relu` serves this call (else the caller runs the torch modules)"""
    return (isinstance(bn, nn.BatchNorm2d) and bn.training and bn.affine and bn.momentum is not None and bn.track_running_stats
            and isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.numel() > 0)


def bn_relu(x, bn, relu=True, conv_bias=None):
    """relu(bn(x)) for a training-mode nn.BatchNorm2d `bn` (its running statistics and num_batches_tracked are updated).
    `conv_bias`: the bias of the convolution that produced `x` WITHOUT adding it — bn(x + b) == bn(x) in training mode, so only the
    running mean needs it and its gradient is exactly zero."""
    if not takes(x, bn):
        raise NotImplementedError("bn_relu takes training-mode affine nn.BatchNorm2d with running statistics on CUDA fp32 NCHW input")
    y = _BnReluTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu, conv_bias)
    with torch.no_grad():
        bn.num_batches_tracked += 1
    return y


class _MaxPoolNCHW(torch.autograd.Function):
    """nn.MaxPool2d whose backward is `gssd_maxpool_nchw_bwd` / `gssd_maxpool_nhwc_bwd` (forward: torch's own kernel with indices)."""

    @staticmethod
    def forward(ctx, x, k, s, pad, ceil_mode):
        cl = _is_channels_last(x) and x.shape[1] % 4 == 0
        y, idx = torch.nn.functional.max_pool2d(x, k, s, pad, 1, ceil_mode, return_indices=True)
        idx = idx.contiguous(memory_format=torch.channels_last) if cl else idx.contiguous()
        ctx.save_for_backward(idx)
        ctx.geom = (x.shape, k, s, pad, cl)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        lib = _lib.require_cuda()
        (idx,) = ctx.saved_tensors
        shape, k, s, pad, cl = ctx.geom
        dev = idx.device
        dyc = dy.detach().to(device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            if cl:
                dyc = dyc.contiguous(memory_format=torch.channels_last)
                dx = torch.empty(shape, dtype=torch.float32, device=dev, memory_format=torch.channels_last)
                _lib.check(lib.gssd_maxpool_nhwc_bwd(dyc.data_ptr(), idx.data_ptr(), shape[0], shape[1], shape[2], shape[3], idx.shape[2],
                                                     idx.shape[3], k, s, pad, dx.data_ptr(), _lib.stream()), "gssd_maxpool_nhwc_bwd")
            else:
                dyc = dyc.contiguous()
                dx = torch.empty(shape, dtype=torch.float32, device=dev)
                _lib.check(lib.gssd_maxpool_nchw_bwd(dyc.data_ptr(), idx.data_ptr(), shape[0] * shape[1], shape[2], shape[3], idx.shape[2],
                                                     idx.shape[3], k, s, pad, dx.data_ptr(), _lib.stream()), "gssd_maxpool_nchw_bwd")
        return dx, None, None, None, None


def _pool_geometry(m):
    """(kernel, stride, padding) of a square, undilated nn.MaxPool2d, else None"""
    def one(v):
        v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        return v[0] if v[0] == v[1] else None
    k, s, p, d = one(m.kernel_size), one(m.stride if m.stride is not None else m.kernel_size), one(m.padding), one(m.dilation)
    return None if None in (k, s, p) or d != 1 or m.return_indices else (int(k), int(s), int(p))


def max_pool(x, m):
    """m(x) for an nn.MaxPool2d `m`, with the gather backward when a gradient will flow (CUDA fp32 NCHW), else the module itself"""
    geom = _pool_geometry(m) if isinstance(m, nn.MaxPool2d) else None
    if (geom is None or not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4)
            or not (torch.is_grad_enabled() and x.requires_grad)):
        return m(x)
    xin = x if _is_channels_last(x) and x.shape[1] % 4 == 0 else x.contiguous()
    return _MaxPoolNCHW.apply(xin, geom[0], geom[1], geom[2], bool(m.ceil_mode))


def to_channels_last(modules):
    """store the filters of every nn.Conv2d of `modules` channels-last, once (values, names and state dict are unchanged): with a
    channels-last input cuDNN then runs its NHWC kernels without converting the filter at every call"""
    for m in modules:
        if isinstance(m, nn.Conv2d) and m.weight.dim() == 4 and not _is_channels_last(m.weight) and m.weight.shape[1] > 1:
            with torch.no_grad():
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)


def run_layers(modules, x, start=0, stop=None, skip_bias=True, tc=None):
    """x through modules[start:stop]; every training-mode (BatchNorm2d, ReLU) pair runs as one fused node, a BatchNorm2d without a
    ReLU behind it as the same kernels without the clamp, a convolution in front of such a BatchNorm2d without its bias add (see
    bn_relu), an nn.MaxPool2d with the gather backward, everything else as the module itself.
    tc: a dict (the caller's cache of packed layers) turns on the tcgen05 path for the grouped backbone convolutions: every run of
    (Conv2d, training-mode BatchNorm2d, ReLU) triples with 64 / 128 channels per group (conv3_2 .. conv5_3 of the reference's vgg
    list) runs in bf16 on the PM layout, forward and backward (source_block.PMConvLayer)."""
    stop = len(modules) if stop is None else stop
    k = start
    while k < stop:
        m = modules[k]
        nxt = modules[k + 1] if k + 1 < stop else None
        if tc is not None and isinstance(m, nn.Conv2d) and isinstance(nxt, nn.BatchNorm2d):
            from .source_block import pm_layers_at, run_pm_layers
            pairs = pm_layers_at(modules, k, stop, x)
            if pairs:
                x = run_pm_layers(pairs, x, tc)
                k += 3 * len(pairs)
                continue
        if (skip_bias and isinstance(m, nn.Conv2d) and m.bias is not None and m.padding_mode == "zeros" and isinstance(nxt, nn.BatchNorm2d)
                and isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
            # convolution -> training-mode BatchNorm: torch's separate bias-add pass behind the cuDNN convolution and the reduction over
            # dy for the bias gradient are skipped (0.25 - 0.9 ms per backbone layer at batch 32); the bias only enters the running mean
            x0 = torch.nn.functional.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)
            if takes(x0, nxt):
                fuse = k + 2 < stop and isinstance(modules[k + 2], nn.ReLU)
                x = bn_relu(x0, nxt, relu=fuse, conv_bias=m.bias)
                k += 3 if fuse else 2
            else:
                x = x0 + m.bias.view(1, -1, 1, 1)
                k += 1
            continue
        if takes(x, m):
            fuse = k + 1 < stop and isinstance(modules[k + 1], nn.ReLU)
            x = bn_relu(x, m, relu=fuse)
            k += 2 if fuse else 1
        elif isinstance(m, nn.MaxPool2d):
            x = max_pool(x, m)
            k += 1
        else:
            x = m(x)
            k += 1
    return x
