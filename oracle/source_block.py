"""numpy restatement of one GSSD source block — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/ssd_liverdet/models/ssd_multiphase_custom_group.py:
  * grouped conv -> BN -> ReLU                      forward, lines 258-259 (vgg[30..32]) / 300-301 (vgg[47..49])
  * L2Norm (source 1 only)                          line 281  -> layers/modules/l2norm.py:19-23
  * fuse_X1 -> bn_fuse_X1 -> ReLU                   lines 290-297 / 317-323 / 365-369
  * loc.k / conf.k 3x3, permute(0,2,3,1), flatten   lines 375-380
nn.Conv2d / nn.BatchNorm2d semantics are torch's (cross-correlation, biased variance for the
normalisation in training mode).  Pinned against the reference's own modules by
tests/golden/make_golden_block.py -> tests/golden/source_block.npz.

`emulate_bf16=True` rounds the operands of every convolution (activations and weights) to bfloat16 the
way the CUDA path stores them, keeping fp32 accumulation: the tight comparison for the kernels; the
plain fp32 result is the reference-faithful one (north-star tolerance 1e-2 relative for the bf16 conv).
"""
import numpy as np


def bf16_round(a):
    """round-to-nearest-even to bfloat16, returned as float32."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    out = ((u + r) & 0xFFFF0000).astype(np.uint32)
    return out.view(np.float32)


def conv2d(x, w, b, groups=1, pad=0):
    """x[N,C,H,W], w[Co,C/groups,kh,kw], stride 1 -> [N,Co,H+2p-kh+1,W+2p-kw+1] (fp32 accumulate in fp64 blocks)"""
    N, C, H, W = x.shape
    Co, Cg, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    Ho, Wo = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    out = np.zeros((N, Co, Ho, Wo), np.float64)
    ng = Co // groups
    for g in range(groups):
        xg = xp[:, g * Cg:(g + 1) * Cg]
        wg = w[g * ng:(g + 1) * ng].astype(np.float64)
        for ky in range(kh):
            for kx in range(kw):
                patch = xg[:, :, ky:ky + Ho, kx:kx + Wo].astype(np.float64)          # [N,Cg,Ho,Wo]
                out[:, g * ng:(g + 1) * ng] += np.einsum("nchw,oc->nohw", patch, wg[:, :, ky, kx], optimize=True)
    if b is not None:
        out += b.astype(np.float64)[None, :, None, None]
    return out.astype(np.float32)


def batch_norm(x, gamma, beta, mean, var, eps, training):
    """F.batch_norm: training -> batch statistics (biased variance); returns (y, batch_mean, batch_var_unbiased)"""
    if training:
        m = x.astype(np.float64).mean(axis=(0, 2, 3))
        v = x.astype(np.float64).var(axis=(0, 2, 3))
        n = x.shape[0] * x.shape[2] * x.shape[3]
        stats = (m.astype(np.float32), (v * n / max(n - 1, 1)).astype(np.float32))
    else:
        m, v = mean.astype(np.float64), var.astype(np.float64)
        stats = None
    y = (x - m[None, :, None, None]) / np.sqrt(v[None, :, None, None] + eps) * gamma[None, :, None, None] + beta[None, :, None, None]
    return y.astype(np.float32), stats


def l2norm(x, weight, eps=1e-10):
    """l2norm.py:19-23"""
    norm = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)) + eps
    return (weight[None, :, None, None] * (x / norm)).astype(np.float32)


def source_block(x, prm, training=False, emulate_bf16=False):
    """x[N,C,H,W] fp32 = input of the grouped conv (or, when prm has no 'gconv_w', the post-ReLU source itself).
    prm: dict of numpy arrays named after the reference's parameters:
       gconv_w/gconv_b/groups/gconv_pad, bn_{w,b,mean,var} (optional), l2norm_w (optional),
       fuse_w/fuse_b, bn_fuse_{w,b,mean,var} (optional), loc_w/loc_b, conf_w/conf_b, bn_eps
    -> dict(x_out, source, loc[N, H*W*A*4], conf[N, H*W*A*ncls])"""
    rd = bf16_round if emulate_bf16 else (lambda a: a)
    eps = float(prm.get("bn_eps", 1e-5))
    if "gconv_w" in prm:
        y = conv2d(rd(x), rd(prm["gconv_w"]), prm["gconv_b"], int(prm["groups"]), int(prm.get("gconv_pad", 1)))
        if "bn_w" in prm:
            if emulate_bf16 and training:
                _, st = batch_norm(y, prm["bn_w"], prm["bn_b"], None, None, eps, True)
                n = y.shape[0] * y.shape[2] * y.shape[3]
                y, _ = batch_norm(rd(y), prm["bn_w"], prm["bn_b"], st[0], st[1] * (n - 1) / n, eps, False)
            else:
                y, _ = batch_norm(y, prm["bn_w"], prm["bn_b"], prm.get("bn_mean"), prm.get("bn_var"), eps, training)
        x_out = np.maximum(y, 0)
    else:
        x_out = x
    x_out = rd(x_out)
    s = x_out
    fuse_w = prm["fuse_w"]
    if "l2norm_w" in prm:
        if emulate_bf16:
            # the CUDA path folds L2Norm.weight into the fuse weights and applies 1/(norm+eps) after the GEMM
            fuse_w = fuse_w * prm["l2norm_w"][None, :, None, None]
            norm = np.sqrt((x_out.astype(np.float64) ** 2).sum(axis=1, keepdims=True)) + 1e-10
            z = conv2d(s, rd(fuse_w), None, 1, 0) / norm + prm["fuse_b"][None, :, None, None]
            z = z.astype(np.float32)
        else:
            s = l2norm(x_out, prm["l2norm_w"])
            z = conv2d(s, fuse_w, prm["fuse_b"], 1, 0)
    else:
        z = conv2d(s, rd(fuse_w), prm["fuse_b"], 1, 0)
    if "bn_fuse_w" in prm:
        if emulate_bf16 and training:
            _, st = batch_norm(z, prm["bn_fuse_w"], prm["bn_fuse_b"], None, None, eps, True)
            n = z.shape[0] * z.shape[2] * z.shape[3]
            z, _ = batch_norm(rd(z), prm["bn_fuse_w"], prm["bn_fuse_b"], st[0], st[1] * (n - 1) / n, eps, False)
        else:
            z, _ = batch_norm(z, prm["bn_fuse_w"], prm["bn_fuse_b"], prm.get("bn_fuse_mean"), prm.get("bn_fuse_var"), eps, training)
    src = rd(np.maximum(z, 0))
    loc = conv2d(src, rd(prm["loc_w"]), prm["loc_b"], 1, 1)
    conf = conv2d(src, rd(prm["conf_w"]), prm["conf_b"], 1, 1)
    N = x.shape[0]
    return dict(x_out=x_out, source=src,
                loc=np.ascontiguousarray(loc.transpose(0, 2, 3, 1)).reshape(N, -1),
                conf=np.ascontiguousarray(conf.transpose(0, 2, 3, 1)).reshape(N, -1))


# ---- backward of the block (what autograd computes on the reference's modules) ---------------------------------------
# Oracle for the conv backward planned in DESIGN.md §7; float64 throughout.  Pinned against autograd on the reference's own
# modules by tests/golden/make_golden_block_bwd.py -> tests/golden/source_block_bwd.npz.

def conv2d_backward(x, w, dy, groups=1, pad=0):
    """gradients of conv2d(x, w, b, groups, pad) (stride 1) w.r.t. (x, w, b) for the upstream gradient dy[N,Co,Ho,Wo]"""
    x, w, dy = x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64)
    N, C, H, W = x.shape
    Co, Cg, kh, kw = w.shape
    ng = Co // groups
    Ho, Wo = dy.shape[2], dy.shape[3]
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(w)
    for g in range(groups):
        dyg = dy[:, g * ng:(g + 1) * ng]
        for ky in range(kh):
            for kx in range(kw):
                patch = xp[:, g * Cg:(g + 1) * Cg, ky:ky + Ho, kx:kx + Wo]
                dw[g * ng:(g + 1) * ng, :, ky, kx] = np.einsum("nohw,nchw->oc", dyg, patch, optimize=True)
                dxp[:, g * Cg:(g + 1) * Cg, ky:ky + Ho, kx:kx + Wo] += np.einsum("nohw,oc->nchw", dyg, w[g * ng:(g + 1) * ng, :, ky, kx],
                                                                                  optimize=True)
    dx = dxp[:, :, pad:pad + H, pad:pad + W] if pad else dxp
    return dx, dw, dy.sum(axis=(0, 2, 3))


def batch_norm_backward(x, gamma, mean, var, eps, training, dy):
    """gradients of batch_norm w.r.t. (x, gamma, beta); training: statistics are functions of x"""
    x, dy, gamma = x.astype(np.float64), dy.astype(np.float64), gamma.astype(np.float64)
    ax = (0, 2, 3)
    if training:
        m, v = x.mean(axis=ax), x.var(axis=ax)
    else:
        m, v = mean.astype(np.float64), var.astype(np.float64)
    inv = 1.0 / np.sqrt(v + eps)
    xh = (x - m[None, :, None, None]) * inv[None, :, None, None]
    dgamma, dbeta = (dy * xh).sum(axis=ax), dy.sum(axis=ax)
    g = dy * gamma[None, :, None, None]
    if training:
        n = x.shape[0] * x.shape[2] * x.shape[3]
        dx = (g - g.sum(axis=ax)[None, :, None, None] / n - xh * (g * xh).sum(axis=ax)[None, :, None, None] / n) * inv[None, :, None, None]
    else:
        dx = g * inv[None, :, None, None]
    return dx, dgamma, dbeta


def l2norm_backward(x, weight, dy, eps=1e-10):
    """gradients of l2norm (l2norm.py:19-23) w.r.t. (x, weight)"""
    x, dy, weight = x.astype(np.float64), dy.astype(np.float64), weight.astype(np.float64)
    n = np.sqrt((x ** 2).sum(axis=1, keepdims=True))
    d = n + eps
    gw = dy * weight[None, :, None, None]
    dweight = (dy * x / d).sum(axis=(0, 2, 3))
    # d(n)/dx = x / n (0 where the pixel is all zero)
    safe_n = np.where(n > 0, n, 1.0)
    dx = gw / d - x / safe_n * ((gw * x).sum(axis=1, keepdims=True) / d ** 2)
    return dx, dweight


def source_block_backward(x, prm, d_loc, d_conf, training=False):
    """Backward of source_block (fp32-faithful path, no bf16 emulation) for upstream gradients d_loc[N, H*W*A*4] and
    d_conf[N, H*W*A*ncls] -> dict of gradients: 'x' and one entry per parameter name of `prm` (gconv_w, gconv_b, bn_w, bn_b,
    l2norm_w, fuse_w, fuse_b, bn_fuse_w, bn_fuse_b, loc_w, loc_b, conf_w, conf_b)."""
    f64 = lambda a: np.asarray(a, np.float64)
    eps = float(prm.get("bn_eps", 1e-5))
    N = x.shape[0]
    # ---- forward, keeping what the backward needs ----
    if "gconv_w" in prm:
        groups, pad = int(prm["groups"]), int(prm.get("gconv_pad", 1))
        y0 = _conv64(x, prm["gconv_w"], prm["gconv_b"], groups, pad)
        y1 = _bn64(y0, prm, "bn", eps, training) if "bn_w" in prm else y0
        x_out = np.maximum(y1, 0)
    else:
        x_out = f64(x)
    s = _l2norm64(x_out, prm["l2norm_w"]) if "l2norm_w" in prm else x_out
    z0 = _conv64(s, prm["fuse_w"], prm["fuse_b"], 1, 0)
    z1 = _bn64(z0, prm, "bn_fuse", eps, training) if "bn_fuse_w" in prm else z0
    src = np.maximum(z1, 0)
    H, W = src.shape[2], src.shape[3]
    # ---- backward ----
    out = {}
    dl = f64(d_loc).reshape(N, H, W, -1).transpose(0, 3, 1, 2)
    dc = f64(d_conf).reshape(N, H, W, -1).transpose(0, 3, 1, 2)
    dsrc_l, out["loc_w"], out["loc_b"] = conv2d_backward(src, prm["loc_w"], dl, 1, 1)
    dsrc_c, out["conf_w"], out["conf_b"] = conv2d_backward(src, prm["conf_w"], dc, 1, 1)
    dz1 = (dsrc_l + dsrc_c) * (z1 > 0)
    if "bn_fuse_w" in prm:
        dz0, out["bn_fuse_w"], out["bn_fuse_b"] = batch_norm_backward(z0, prm["bn_fuse_w"], prm.get("bn_fuse_mean"), prm.get("bn_fuse_var"),
                                                                       eps, training, dz1)
    else:
        dz0 = dz1
    ds, out["fuse_w"], out["fuse_b"] = conv2d_backward(s, prm["fuse_w"], dz0, 1, 0)
    if "l2norm_w" in prm:
        dx_out, out["l2norm_w"] = l2norm_backward(x_out, prm["l2norm_w"], ds)
    else:
        dx_out = ds
    if "gconv_w" in prm:
        dy1 = dx_out * (y1 > 0)
        if "bn_w" in prm:
            dy0, out["bn_w"], out["bn_b"] = batch_norm_backward(y0, prm["bn_w"], prm.get("bn_mean"), prm.get("bn_var"), eps, training, dy1)
        else:
            dy0 = dy1
        out["x"], out["gconv_w"], out["gconv_b"] = conv2d_backward(x, prm["gconv_w"], dy0, groups, pad)
    else:
        out["x"] = dx_out
    return out


def _conv64(x, w, b, groups, pad):
    """conv2d in float64 (the float32 conv2d above rounds its result)"""
    x, w = np.asarray(x, np.float64), np.asarray(w, np.float64)
    N, C, H, W = x.shape
    Co, Cg, kh, kw = w.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    Ho, Wo = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    out = np.zeros((N, Co, Ho, Wo), np.float64)
    ng = Co // groups
    for g in range(groups):
        for ky in range(kh):
            for kx in range(kw):
                out[:, g * ng:(g + 1) * ng] += np.einsum("nchw,oc->nohw", xp[:, g * Cg:(g + 1) * Cg, ky:ky + Ho, kx:kx + Wo],
                                                         w[g * ng:(g + 1) * ng, :, ky, kx], optimize=True)
    return out + np.asarray(b, np.float64)[None, :, None, None]


def _bn64(x, prm, name, eps, training):
    if training:
        m, v = x.mean(axis=(0, 2, 3)), x.var(axis=(0, 2, 3))
    else:
        m, v = np.asarray(prm[name + "_mean"], np.float64), np.asarray(prm[name + "_var"], np.float64)
    g, b = np.asarray(prm[name + "_w"], np.float64), np.asarray(prm[name + "_b"], np.float64)
    return (x - m[None, :, None, None]) / np.sqrt(v[None, :, None, None] + eps) * g[None, :, None, None] + b[None, :, None, None]


def _l2norm64(x, weight, eps=1e-10):
    n = np.sqrt((x ** 2).sum(axis=1, keepdims=True)) + eps
    return np.asarray(weight, np.float64)[None, :, None, None] * (x / n)


# ---- the grouped backbone triples conv3_2 .. conv5_3 (models/ssd_multiphase_custom_group.py:434-460: nn.Conv2d(groups=4),
# nn.BatchNorm2d, nn.ReLU; run by GSSD:254-259 / 300-301 as `x = self.vgg[k](x)`) — forward and backward in training mode ---------
def backbone_triples(x, prm, gout=None, groups=4, eps=1e-5, momentum=0.1):
    """x[N,C,H,W]; prm: list of dicts (w, b, gamma, beta) of consecutive [conv 3x3 pad 1, BatchNorm2d (batch statistics), ReLU]
    triples -> dict(y, running statistics after one step from (0, 1)); with gout (upstream gradient of y) also
    x / conv_w / conv_b / bn_w / bn_b gradients, keyed as tests/golden/backbone_bwd.npz.  Pinned against autograd on the reference's
    own modules in float64 by tests/test_oracle_golden.py."""
    h = np.asarray(x, np.float64)
    saved, out = [], {}
    for t, p in enumerate(prm):
        w, b = p["w"].astype(np.float64), p["b"].astype(np.float64)
        z = _conv64(h, w, b, groups, 1)
        mean, var = z.mean(axis=(0, 2, 3)), z.var(axis=(0, 2, 3))
        n = z.shape[0] * z.shape[2] * z.shape[3]
        out["%d.running_mean" % t] = momentum * mean
        out["%d.running_var" % t] = (1 - momentum) * 1.0 + momentum * var * n / max(n - 1, 1)
        a = (z - mean[None, :, None, None]) / np.sqrt(var[None, :, None, None] + eps) * p["gamma"].astype(np.float64)[None, :, None, None] \
            + p["beta"].astype(np.float64)[None, :, None, None]
        saved.append((h, w, z, mean, var, a))
        h = np.maximum(a, 0)
    out["y"] = h
    if gout is None:
        return out
    d = np.asarray(gout, np.float64)
    for t in range(len(prm) - 1, -1, -1):
        hin, w, z, mean, var, a = saved[t]
        d = d * (a > 0)
        dz, dgamma, dbeta = batch_norm_backward(z, prm[t]["gamma"], mean, var, eps, True, d)
        d, dw, db = conv2d_backward(hin, w, dz, groups, 1)
        out.update({"%d.conv_w" % t: dw, "%d.conv_b" % t: db, "%d.bn_w" % t: dgamma, "%d.bn_b" % t: dbeta})
    out["x"] = d
    return out
