"""L2Norm — same module as layers/modules/l2norm.py:7-23 of the reference (parameter name `weight`,
constant init), forward and backward as libgssd_b200.so kernels."""
import torch
import torch.nn as nn
import torch.nn.init as init

from ... import _lib


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps):
        lib = _lib.require_cuda()
        dev = _lib.device_of(x, weight)
        with torch.cuda.device(dev):
            xc, w = _lib.f32(x, dev), _lib.f32(weight, dev)
            B, Cn = xc.size(0), xc.size(1)
            HW = xc.numel() // (B * Cn)
            y = torch.empty_like(xc)
            norm = torch.empty((B, HW), dtype=torch.float32, device=dev)
            _lib.check(lib.gssd_l2norm_fwd(xc.data_ptr(), w.data_ptr(), B, Cn, HW, float(eps), y.data_ptr(),
                                           norm.data_ptr(), _lib.stream()), "gssd_l2norm_fwd")
        ctx.save_for_backward(xc, w, norm)
        ctx.eps = eps
        ctx.src = (x.device, weight.device)
        return y.to(x.device)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        xc, w, norm = ctx.saved_tensors
        dev = xc.device
        with torch.cuda.device(dev):
            g = _lib.f32(gy, dev)
            B, Cn = xc.size(0), xc.size(1)
            HW = xc.numel() // (B * Cn)
            gx = torch.empty_like(xc)
            gw = torch.empty_like(w)
            nb = lib.gssd_l2norm_bwd_ws_bytes(B, Cn, HW)
            ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
            _lib.check(lib.gssd_l2norm_bwd(xc.data_ptr(), w.data_ptr(), norm.data_ptr(), g.data_ptr(), B, Cn, HW,
                                           float(ctx.eps), gx.data_ptr(), gw.data_ptr(), ws.data_ptr(), nb,
                                           _lib.stream()), "gssd_l2norm_bwd")
        return gx.to(ctx.src[0]), gw.to(ctx.src[1]), None


class L2Norm(nn.Module):
    def __init__(self, n_channels, scale):
        super(L2Norm, self).__init__()
        self.n_channels = n_channels
        self.gamma = scale or None
        self.eps = 1e-10
        self.weight = nn.Parameter(torch.Tensor(self.n_channels))
        self.reset_parameters()

    def reset_parameters(self):
        init.constant_(self.weight, self.gamma)

    def forward(self, x):
        """x[B,C,H,W] -> weight[c] * x / (sqrt(sum_c x^2) + eps)   [l2norm.py:19-23]"""
        return _L2NormFn.apply(x, self.weight, self.eps)
