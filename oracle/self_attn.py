"""numpy restatement of GSSD++'s Self_Attn block — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/ssd_liverdet/layers/self_attn.py:47-89 (forward) with the spectrally normalised 1x1 convolutions of
layers/spectral_norm.py in evaluation mode (no power iteration: weight = weight_orig / (u^T W v), spectral_norm.py:69-95) and
torch's adaptive_avg_pool2d windows (start = floor(i*H/out), end = ceil((i+1)*H/out)).  `attention` / `attention_backward`
are the core the CUDA kernels replace (self_attn.py:69-81).  Pinned against the reference's own module by
tests/golden/make_golden_self_attn.py -> tests/golden/self_attn.npz (tests/test_oracle_golden.py).  float64."""
import numpy as np


def attention(theta, phi, g):
    """theta [B,D,N], phi [B,D,M], g [B,Cv,M] -> attn [B,N,M], attn_g [B,Cv,N]   (self_attn.py:71-72, 80)"""
    theta, phi, g = (np.asarray(a, np.float64) for a in (theta, phi, g))
    s = np.einsum("bdn,bdm->bnm", theta, phi)
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    attn = e / e.sum(-1, keepdims=True)
    return attn, np.einsum("bcm,bnm->bcn", g, attn)


def attention_backward(theta, phi, g, attn, d_attn_g):
    theta, phi, g, attn, d_o = (np.asarray(a, np.float64) for a in (theta, phi, g, attn, d_attn_g))
    d_g = np.einsum("bcn,bnm->bcm", d_o, attn)
    d_p = np.einsum("bcn,bcm->bnm", d_o, g)
    d_s = attn * (d_p - (d_p * attn).sum(-1, keepdims=True))
    return np.einsum("bnm,bdm->bdn", d_s, phi), np.einsum("bnm,bdn->bdm", d_s, theta), d_g


def adaptive_avg_pool(x, out):
    B, C, H, W = x.shape
    y = np.zeros((B, C, out, out))
    for i in range(out):
        y0, y1 = (i * H) // out, -((-(i + 1) * H) // out)
        for j in range(out):
            x0, x1 = (j * W) // out, -((-(j + 1) * W) // out)
            y[:, :, i, j] = x[:, :, y0:y1, x0:x1].mean((2, 3))
    return y


def adaptive_avg_pool_backward(dy, H, W):
    B, C, out, _ = dy.shape
    dx = np.zeros((B, C, H, W))
    for i in range(out):
        y0, y1 = (i * H) // out, -((-(i + 1) * H) // out)
        for j in range(out):
            x0, x1 = (j * W) // out, -((-(j + 1) * W) // out)
            dx[:, :, y0:y1, x0:x1] += dy[:, :, i:i + 1, j:j + 1] / ((y1 - y0) * (x1 - x0))
    return dx


def sn_weight(prm, name):
    w = np.asarray(prm[name + ".weight_orig"], np.float64)
    mat = w.reshape(w.shape[0], -1)
    sigma = np.asarray(prm[name + ".weight_u"], np.float64) @ (mat @ np.asarray(prm[name + ".weight_v"], np.float64))
    return mat / sigma, np.asarray(prm[name + ".bias"], np.float64)


def conv1x1(x, wb):
    w, b = wb
    return np.einsum("oc,bchw->bohw", w, x) + b[None, :, None, None]


def self_attn_forward(x, prm, max_pool_factor=1):
    """-> (out, sigma*attn_g, attn) as Self_Attn.forward(x, True) in eval mode"""
    x = np.asarray(x, np.float64)
    B, C, H, W = x.shape
    pooled = max(H // max_pool_factor, 1)
    theta = conv1x1(x, sn_weight(prm, "snconv1x1_theta")).reshape(B, C // 8, H * W)
    phi = adaptive_avg_pool(conv1x1(x, sn_weight(prm, "snconv1x1_phi")), pooled).reshape(B, C // 8, -1)
    g = adaptive_avg_pool(conv1x1(x, sn_weight(prm, "snconv1x1_g")), pooled).reshape(B, C // 2, -1)
    attn, attn_g = attention(theta, phi, g)
    gated = float(np.asarray(prm["sigma"]).reshape(-1)[0]) * conv1x1(attn_g.reshape(B, C // 2, H, W), sn_weight(prm, "snconv1x1_attn"))
    return x + gated, gated, attn, (theta, phi, g, attn_g)
