"""Drop-in for the part of the reference's `layers/self_attn.py` that GSSD++ uses (models/ssd_multiphase_custom_group.py:142-153,
261-288): the SAGAN-style `Self_Attn` block with its spectrally normalised 1x1 convolutions, and the `snconv2d` / `snlinear` /
`sn_embedding` / `init_weights` helpers of the same file (self_attn.py:10-27).  The unused GAN generator / discriminator of
that file are not part of the detector and are not provided.

The attention core — scores, softmax and the product with `g` (self_attn.py:69-81) — runs as ONE kernel of libgssd_b200.so
forward (`gssd_attn_fwd`) and two backward (`gssd_attn_bwd`), in fp32; the 1x1 convolutions, the average pooling and the
residual stay torch operators on the module's own parameters.  Parameter and buffer names are the reference's
(`sigma`, `snconv1x1_{theta,phi,g,attn}.{bias,weight_orig,weight_u,weight_v}`): its spectral_norm.py is torch's
`torch.nn.utils.spectral_norm`, which is what is used here, so checkpoints load both ways.  No CPU fallback."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

from .. import _lib


def init_weights(m):
    """self_attn.py:10-13"""
    if type(m) in (nn.Linear, nn.Conv2d):
        nn.init.xavier_uniform_(m.weight)
        m.bias.data.fill_(0.)


def snconv2d(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
    return spectral_norm(nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                                   groups=groups, bias=bias))


def snlinear(in_features, out_features):
    return spectral_norm(nn.Linear(in_features, out_features))


def sn_embedding(num_embeddings, embedding_dim):
    return spectral_norm(nn.Embedding(num_embeddings, embedding_dim))


class _AttentionCore(torch.autograd.Function):
    """(theta [B,D,N], phi [B,D,M], g [B,Cv,M]) -> (attn_g [B,Cv,N], attn [B,N,M]); attn carries no gradient (it is returned for
    inspection only, self_attn.py:86)."""

    @staticmethod
    def forward(ctx, theta, phi, g):
        lib = _lib.require_cuda()
        if not theta.is_cuda:
            raise RuntimeError("gssd_attn needs CUDA tensors; there is no CPU fallback")
        dev = theta.device
        theta, phi, g = _lib.f32(theta, dev), _lib.f32(phi, dev), _lib.f32(g, dev)
        B, D, N = theta.shape
        M, Cv = phi.shape[2], g.shape[1]
        if tuple(phi.shape) != (B, D, M) or tuple(g.shape) != (B, Cv, M):
            raise ValueError("theta [B,D,N], phi [B,D,M], g [B,Cv,M] expected, got %s %s %s" % (tuple(theta.shape), tuple(phi.shape), tuple(g.shape)))
        attn = torch.empty((B, N, M), dtype=torch.float32, device=dev)
        out = torch.empty((B, Cv, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.gssd_attn_fwd(theta.data_ptr(), phi.data_ptr(), g.data_ptr(), B, D, Cv, N, M, attn.data_ptr(), out.data_ptr(), _lib.stream())
        if rc == _lib.ERR_LIMIT:
            raise NotImplementedError("gssd_attn: C/8 and C/2 must be multiples of 32 and a strip of 8 queries against all keys must "
                                      "fit in shared memory (got D %d, Cv %d, N %d, M %d)" % (D, Cv, N, M))
        _lib.check(rc, "gssd_attn_fwd")
        ctx.save_for_backward(theta, phi, g, attn)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out, _d_attn):
        lib = _lib.require_cuda()
        theta, phi, g, attn = ctx.saved_tensors
        dev = theta.device
        B, D, N = theta.shape
        M, Cv = phi.shape[2], g.shape[1]
        d_out = _lib.f32(d_out, dev)
        d_theta, d_phi, d_g = torch.empty_like(theta), torch.empty_like(phi), torch.empty_like(g)
        ws = torch.empty((B, N, M), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.gssd_attn_bwd(theta.data_ptr(), phi.data_ptr(), g.data_ptr(), attn.data_ptr(), d_out.data_ptr(), B, D, Cv, N, M,
                                         d_theta.data_ptr(), d_phi.data_ptr(), d_g.data_ptr(), ws.data_ptr(), _lib.stream()), "gssd_attn_bwd")
        return d_theta, d_phi, d_g


attention_core = _AttentionCore.apply


class Self_Attn(nn.Module):
    """self_attn.py:29-89.  `forward(x, return_attn_map)` -> (x + sigma*attn_g, sigma*attn_g[, attn])."""

    def __init__(self, in_channels, max_pool_factor=1):
        super().__init__()
        self.in_channels = in_channels
        self.snconv1x1_theta = snconv2d(in_channels, in_channels // 8, kernel_size=1)
        self.snconv1x1_phi = snconv2d(in_channels, in_channels // 8, kernel_size=1)
        self.snconv1x1_g = snconv2d(in_channels, in_channels // 2, kernel_size=1)
        self.snconv1x1_attn = snconv2d(in_channels // 2, in_channels, kernel_size=1)
        self.softmax = nn.Softmax(dim=-1)                    # kept for state / repr parity; the kernel computes it
        self.sigma = nn.Parameter(torch.zeros(1))
        self.max_pool_factor = max_pool_factor

    def forward(self, x, return_attn_map=False):
        b, ch, h, w = x.shape
        assert h == w
        pooled = max(int(h // self.max_pool_factor), 1)     # keys live on a (pooled x pooled) grid (self_attn.py:57-59)
        theta = self.snconv1x1_theta(x).reshape(b, ch // 8, h * w)
        phi = F.adaptive_avg_pool2d(self.snconv1x1_phi(x), pooled).reshape(b, ch // 8, pooled * pooled)
        g = F.adaptive_avg_pool2d(self.snconv1x1_g(x), pooled).reshape(b, ch // 2, pooled * pooled)
        attn_g, attn = attention_core(theta, phi, g)
        gated = self.sigma * self.snconv1x1_attn(attn_g.reshape(b, ch // 2, h, w))
        out = x + gated
        return (out, gated, attn) if return_attn_map else (out, gated)
