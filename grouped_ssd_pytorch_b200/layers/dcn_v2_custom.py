"""Drop-in for the reference's `layers/dcn_v2_custom.py` (GSSD++'s modulated deformable convolution) with the operator itself —
`dcn_v2._DCNv2.apply`, a third-party CUDA extension the reference does not vendor (dcn_v2_custom.py:13) — running on this
library's sm_100a kernels: deformable im2col (`gssd_dcn_columns`) + the tcgen05/TMEM GEMM of the source block
(`gssd_conv_igemm` over the columns) forward; `gssd_conv_igemm` / `gssd_conv_wgrad` / `gssd_dcn_columns_bwd` backward.

Same names, constructor arguments, parameter names (`weight`, `bias`, `conv_offset_mask.*`) and return values as the reference:
`DCNv2.forward(input, offset, mask)` (dcn_v2_custom.py:43-55), `DCN.forward(input) -> (out, offset)` (:79-88),
`dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups)` (:15).

bf16 operands with fp32 accumulation: results agree with an fp32 evaluation of the same operator
(`torchvision.ops.deform_conv2d`, the stand-in SURVEY App. A names) within 1e-2 of each tensor's scale — the north-star
tolerance of the bf16 convolution path.  No CPU fallback; shapes the kernels do not take raise NotImplementedError."""
import math
import weakref

import torch
from torch import nn
from torch.nn.modules.utils import _pair

from .. import _lib
from .modules.source_block import PM, conv_igemm, _wgrad


class _Packed(object):
    """what `conv_igemm` needs to know about a packed filter matrix"""
    __slots__ = ("w", "c_in", "c_out", "groups", "taps")

    def __init__(self, w, c_in, c_out):
        self.w, self.c_in, self.c_out, self.groups, self.taps = w, c_in, c_out, 1, 1


_PACKED = {}          # id(weight tensor) -> (weak reference to it, version, device, packed bf16 filter matrix)


def _packed_filter(weight, dev):
    """bf16 [c_out, 9*c_in], k = tap*c_in + c (the columns' order), re-packed only when the parameter changed"""
    key = id(weight)
    ent = _PACKED.get(key)
    if ent is not None and ent[0]() is weight and ent[1] == weight._version and ent[2] == dev:
        return ent[3]
    c_out, c_in = weight.shape[0], weight.shape[1]
    wf = _lib.f32(weight, dev)
    wp = torch.empty((c_out, 9 * c_in), dtype=torch.bfloat16, device=dev)
    _lib.check(_lib.load().gssd_conv_pack_weights(wf.data_ptr(), c_out, c_in, 1, 9, None, wp.data_ptr(), _lib.stream()),
               "gssd_conv_pack_weights")
    _PACKED[key] = (weakref.ref(weight, lambda _r, k=key: _PACKED.pop(k, None)), weight._version, dev, wp)
    return wp


def _check_geometry(weight, stride, padding, dilation, c_in, deformable_groups):
    kh, kw = weight.shape[2], weight.shape[3]
    if (kh, kw) != (3, 3) or _pair(stride) != (1, 1) or _pair(padding) != (1, 1) or _pair(dilation) != (1, 1):
        raise NotImplementedError("gssd_dcn takes the 3x3 / stride 1 / padding 1 / dilation 1 deformable convolution of GSSD++ "
                                  "(dcn_v2_custom.py:164-173 of the model file), got kernel %dx%d stride %s padding %s dilation %s"
                                  % (kh, kw, stride, padding, dilation))
    if weight.shape[1] != c_in:
        raise ValueError("weight expects %d input channels, input has %d" % (weight.shape[1], c_in))
    if c_in % 128 or weight.shape[0] % 64 or c_in % deformable_groups or (c_in // deformable_groups) % 8:
        raise NotImplementedError("gssd_dcn: c_in must be a multiple of 128, c_out of 64, c_in / deformable_groups of 8 "
                                  "(got c_in %d, c_out %d, deformable_groups %d)" % (c_in, weight.shape[0], deformable_groups))


class _DCNv2(torch.autograd.Function):
    """`dcn_v2._DCNv2` (the operator behind dcn_v2_custom.py:15)."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
        lib = _lib.require_cuda()
        if not input.is_cuda:
            raise RuntimeError("gssd_dcn needs CUDA tensors; there is no CPU fallback")
        n, c_in, h, w = input.shape
        dg = int(deformable_groups)
        _check_geometry(weight, stride, padding, dilation, c_in, dg)
        c_out = weight.shape[0]
        if tuple(offset.shape) != (n, 2 * dg * 9, h, w) or tuple(mask.shape) != (n, dg * 9, h, w):
            raise ValueError("offset / mask must be [N, %d, H, W] / [N, %d, H, W]" % (2 * dg * 9, dg * 9))
        dev = input.device
        with torch.cuda.device(dev):
            st = _lib.stream()
            x = PM.from_nchw(input)
            off, msk = _lib.f32(offset, dev), _lib.f32(mask, dev)
            col = PM.empty(n, 9 * c_in, h, w, dev)
            _lib.check(lib.gssd_dcn_columns(x.data.data_ptr(), off.data_ptr(), msk.data_ptr(), n, c_in, h, w, dg,
                                            col.data.data_ptr(), st), "gssd_dcn_columns")
            wp = _packed_filter(weight, dev)
            b = None if bias is None else _lib.f32(bias, dev)
            y = conv_igemm(col, _Packed(wp, 9 * c_in, c_out), relu=False, shift=b)
            out = y.to_nchw()
        ctx.save_for_backward(off, msk)
        ctx.x, ctx.col, ctx.wp = x, col, wp
        ctx.dg, ctx.has_bias = dg, bias is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _lib.require_cuda()
        off, msk = ctx.saved_tensors
        x, col, wp, dg = ctx.x, ctx.col, ctx.wp, ctx.dg
        n, c_in, h, w = x.n, x.c, x.h, x.w
        c_out = wp.shape[0]
        dev = x.data.device
        need_x, need_off, need_mask, need_w, need_b = ctx.needs_input_grad[:5]
        d_in = d_off = d_mask = d_w = d_b = None
        with torch.cuda.device(dev):
            st = _lib.stream()
            dy = PM.from_nchw(grad_out)
            if need_w:
                d_w = _wgrad(dy, col, c_out, 1, 1).reshape(c_out, 9, c_in).permute(0, 2, 1).reshape(c_out, c_in, 3, 3)
            if need_b and ctx.has_bias:
                d_b = grad_out.sum((0, 2, 3))
            if need_x or need_off or need_mask:
                dcol = conv_igemm(dy, _Packed(wp.t().contiguous(), c_out, 9 * c_in), relu=False)       # d_columns = dY * W
                dx_pm = torch.empty((x.rows, c_in), dtype=torch.float32, device=dev)
                d_off, d_mask = torch.empty_like(off), torch.empty_like(msk)
                _lib.check(lib.gssd_dcn_columns_bwd(x.data.data_ptr(), off.data_ptr(), msk.data_ptr(), dcol.data.data_ptr(), n, c_in,
                                                    h, w, dg, dx_pm.data_ptr(), d_off.data_ptr(), d_mask.data_ptr(), st),
                           "gssd_dcn_columns_bwd")
                if need_x:
                    d_in = torch.empty((n, c_in, h, w), dtype=torch.float32, device=dev)
                    _lib.check(lib.gssd_pmf32_to_nchw(dx_pm.data_ptr(), n, c_in, h, w, d_in.data_ptr(), st), "gssd_pmf32_to_nchw")
        return (d_in, d_off if need_off else None, d_mask if need_mask else None, d_w, d_b, None, None, None, None)


dcn_v2_conv = _DCNv2.apply


class DCNv2(nn.Module):
    """dcn_v2_custom.py:18-55: the deformable convolution with offsets and masks supplied by the caller."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        # dcn_v2_custom.py:35-41: uniform in +-1/sqrt(fan_in), zero bias
        bound = 1.0 / math.sqrt(self.in_channels * self.kernel_size[0] * self.kernel_size[1])
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            self.bias.zero_()

    def forward(self, input, offset, mask):
        taps = self.kernel_size[0] * self.kernel_size[1]
        assert offset.shape[1] == 2 * self.deformable_groups * taps
        assert mask.shape[1] == self.deformable_groups * taps
        return dcn_v2_conv(input, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)


class DCN(DCNv2):
    """dcn_v2_custom.py:58-88: offsets and masks come from a zero-initialised 3x3 convolution of the input; returns
    (output, offset)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        taps = self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(self.in_channels, self.deformable_groups * 3 * taps, kernel_size=self.kernel_size,
                                          stride=self.stride, padding=self.padding, bias=True)
        self.init_offset()

    def init_offset(self):
        with torch.no_grad():
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()

    def forward(self, input):
        o1, o2, mask = torch.chunk(self.conv_offset_mask(input), 3, dim=1)
        offset = torch.cat((o1, o2), dim=1)
        out = dcn_v2_conv(input, offset, torch.sigmoid(mask), self.weight, self.bias, self.stride, self.padding, self.dilation,
                          self.deformable_groups)
        return out, offset
