"""numpy restatement of GSSD++'s modulated deformable convolution (DCNv2) — TEST INFRASTRUCTURE ONLY.

The operator is `dcn_v2._DCNv2.apply`, called at /root/reference/ssd_liverdet/layers/dcn_v2_custom.py:49-55 and 84-88.  It is a
third-party CUDA extension (CharlesShang/DCNv2) that the reference neither vendors nor pins (dcn_v2_custom.py:13 is the only
trace), so what is restated here is its published algorithm — deformable im2col + GEMM:

    col[n, tap, c, y, x] = mask[n, g*9+tap, y, x] * bilinear(x[n, c], y - 1 + ky + off_y, x - 1 + kx + off_x)
    out[n, o, y, x]      = bias[o] + sum_{tap, c} weight[o, c, ky, kx] * col[n, tap, c, y, x]

with g = c // (C/dg), off_y / off_x = offset[n, (g*9+tap)*2 + 0 / 1], a sample that is zero unless -1 < y < H and -1 < x < W and
neighbours outside the image counting as zero.  Parity is pinned against `torchvision.ops.deform_conv2d` (the same operator; the
stand-in SURVEY App. A uses to run GSSD++ here) driven through the reference's OWN `DCN` module: tests/golden/make_golden_dcn.py
-> tests/golden/dcn.npz, checked by tests/test_oracle_golden.py.  The offset gradient follows torchvision's
get_coordinate_weight: each neighbour is validated on its own, without the in-range gate of the forward (DCNv2's own kernel has
the gate; the two differ only for samples at exactly y = -1 or x = -1, which zero offsets do produce on the top / left border).
All arithmetic in float64."""
import numpy as np


def _geometry(offset, n, g, tap, H, W):
    ky, kx = divmod(tap, 3)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    sy = yy - 1 + ky + offset[n, (g * 9 + tap) * 2]
    sx = xx - 1 + kx + offset[n, (g * 9 + tap) * 2 + 1]
    inr = (sy > -1) & (sy < H) & (sx > -1) & (sx < W)
    y0, x0 = np.floor(sy), np.floor(sx)
    ly, lx = sy - y0, sx - x0
    return sy, sx, inr, y0.astype(np.int64), x0.astype(np.int64), ly, lx


def _neighbours(xg, y0, x0):
    """xg [c, H, W] -> the four neighbours [c, H, W] each, zero where the neighbour is not an image pixel; validity masks."""
    _, H, W = xg.shape
    out, valid = [], []
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
        yy, xx = y0 + dy, x0 + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = xg[:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok
        out.append(v)
        valid.append(ok)
    return out, valid


def columns(x, offset, mask, dg):
    """[N, 9, C, H, W] float64"""
    x, offset, mask = (np.asarray(a, np.float64) for a in (x, offset, mask))
    N, C, H, W = x.shape
    cpg = C // dg
    col = np.zeros((N, 9, C, H, W))
    for n in range(N):
        for g in range(dg):
            xg = x[n, g * cpg:(g + 1) * cpg]
            for tap in range(9):
                _, _, inr, y0, x0, ly, lx = _geometry(offset, n, g, tap, H, W)
                (v1, v2, v3, v4), _ = _neighbours(xg, y0, x0)
                val = ((1 - ly) * (1 - lx)) * v1 + ((1 - ly) * lx) * v2 + (ly * (1 - lx)) * v3 + (ly * lx) * v4
                col[n, tap, g * cpg:(g + 1) * cpg] = mask[n, g * 9 + tap] * (val * inr)
    return col


def forward(x, offset, mask, weight, bias, dg):
    col = columns(x, offset, mask, dg)
    w2 = np.asarray(weight, np.float64).reshape(weight.shape[0], weight.shape[1], 9).transpose(0, 2, 1)      # [o, tap, c]
    out = np.einsum("ntchw,otc->nohw", col, w2)
    if bias is not None:
        out += np.asarray(bias, np.float64)[None, :, None, None]
    return out


def backward(x, offset, mask, weight, dg, grad_out, has_bias=True):
    """-> dict(d_input, d_offset, d_mask, d_weight, d_bias)"""
    x, offset, mask, weight, grad_out = (np.asarray(a, np.float64) for a in (x, offset, mask, weight, grad_out))
    N, C, H, W = x.shape
    O = weight.shape[0]
    cpg = C // dg
    w2 = weight.reshape(O, C, 9).transpose(0, 2, 1)
    col = columns(x, offset, mask, dg)
    dcol = np.einsum("nohw,otc->ntchw", grad_out, w2)
    d_w = np.einsum("nohw,ntchw->otc", grad_out, col).transpose(0, 2, 1).reshape(O, C, 3, 3)
    d_x, d_off, d_mask = np.zeros_like(x), np.zeros_like(offset), np.zeros_like(mask)
    for n in range(N):
        for g in range(dg):
            sl = slice(g * cpg, (g + 1) * cpg)
            xg = x[n, sl]
            for tap in range(9):
                _, _, inr, y0, x0, ly, lx = _geometry(offset, n, g, tap, H, W)
                (v1, v2, v3, v4), valid = _neighbours(xg, y0, x0)
                d = dcol[n, tap, sl]
                m = mask[n, g * 9 + tap]
                wts = ((1 - ly) * (1 - lx), (1 - ly) * lx, ly * (1 - lx), ly * lx)
                val = (wts[0] * v1 + wts[1] * v2 + wts[2] * v3 + wts[3] * v4) * inr
                d_mask[n, g * 9 + tap] = (d * val).sum(0)
                d_off[n, (g * 9 + tap) * 2] = m * (d * (lx * (v4 - v2) + (1 - lx) * (v3 - v1))).sum(0)
                d_off[n, (g * 9 + tap) * 2 + 1] = m * (d * (ly * (v4 - v3) + (1 - ly) * (v2 - v1))).sum(0)
                for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                    ok = valid[k] & inr
                    ys, xs = np.nonzero(ok)
                    contrib = d[:, ys, xs] * (m * wts[k])[ys, xs]
                    np.add.at(d_x[n, sl], (slice(None), (y0 + dy)[ys, xs], (x0 + dx)[ys, xs]), contrib)
    return dict(d_input=d_x, d_offset=d_off, d_mask=d_mask, d_weight=d_w, d_bias=grad_out.sum((0, 2, 3)) if has_bias else None)
