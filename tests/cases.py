"""Seeded inputs shared by the golden generator (tests/golden/make_golden.py), the oracle tests and
the GPU parity tests.  Everything here is regenerated from seeds with frozen numpy streams."""
import os

import numpy as np

from grouped_ssd_pytorch_b200 import config as cfg
from grouped_ssd_pytorch_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VAR = (0.1, 0.2)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def small_cfg():
    c = dict(cfg.v2)
    c.update(feature_maps=[5, 3, 1], steps=[60, 100, 300], min_sizes=[60, 150, 240],
             max_sizes=[150, 240, 315], aspect_ratios=[[2], [2, 3], [2]])
    return c


def priors(name):
    """golden prior set by name (reference output)."""
    return golden("priors")[name]


def match_rand_targets():
    return syn.targets(syn.rng(11), 3, 1, 5)


def loss_case(tag):
    """-> (loc, conf, priors, targets, num_classes, ratio)"""
    if tag == "a":
        pri = priors("v2"); r = syn.rng(21)
        tg = syn.targets(r, 4, 1, 5)
        return syn.loc(r, 4, pri.shape[0]), syn.conf_logits(r, 4, pri.shape[0], 2), pri, tg, 2, 3
    if tag == "b":
        pri = priors("v2"); r = syn.rng(22)
        tg = syn.targets(r, 3, 1, 8)
        for t in tg:
            t[:, 4] = r.randint(0, 2, size=t.shape[0])
        return syn.loc(r, 3, pri.shape[0]), syn.conf_logits(r, 3, pri.shape[0], 3) * 2, pri, tg, 3, 2
    if tag == "c":
        pri = priors("small"); r = syn.rng(23)
        tg = syn.targets(r, 2, 60, 64)
        return syn.loc(r, 2, pri.shape[0]), syn.conf_logits(r, 2, pri.shape[0], 2), pri, tg, 2, 4
    if tag == "d":
        pri = priors("v2_512"); r = syn.rng(24)
        tg = syn.targets(r, 2, 1, 32)
        return syn.loc(r, 2, pri.shape[0]), syn.conf_logits(r, 2, pri.shape[0], 2), pri, tg, 2, 3
    raise KeyError(tag)


def detect_case(tag):
    """-> (loc, conf, priors, C, thr)"""
    pri = priors("v2")
    P = pri.shape[0]
    table = {"a": (41, 2, 2, -4.0, 0.5, 0.2), "b": (42, 2, 2, 0.0, 0.05, 0.2), "c": (43, 1, 3, -3.0, 0.2, 0.01),
             "d": (44, 1, 2, -30.0, 0.5, 0.2)}
    seed, B, C, shift, sigma, thr = table[tag]
    r = syn.rng(seed)
    loc = syn.loc(r, B, P, sigma)
    conf = syn.detect_scores(r, B, P, C, shift)
    return loc, conf, pri, C, thr


def dense_grads(g, prefix, shape):
    out = np.zeros(int(np.prod(shape)), np.float32)
    out[g[prefix + "_idx"]] = g[prefix + "_val"]
    return out.reshape(shape)


def ohnm_unambiguous(key, pos, ratio):
    """per image: is the reference's pos|neg set independent of how its unstable sort orders equal keys?
    True when no key tie straddles the num_neg cut, or when every prior tied there is a positive
    (positives are in the union anyway)."""
    B, P = key.shape
    ok = np.zeros(B, bool)
    for b in range(B):
        nn = min(ratio * int(pos[b].sum()), P - 1)
        ks = np.sort(key[b])[::-1]
        if nn <= 0 or nn >= P or ks[nn - 1] != ks[nn]:
            ok[b] = True
        else:
            ok[b] = bool(pos[b][key[b] == ks[nn]].all())
    return ok


# ---- source block -------------------------------------------------------------------------
BLOCK_CASES = {
    # tag: (seed, N, C_in, H, W, grouped conv (C_out, groups, k) or None, bn, l2norm, C_fuse, anchors, classes, training)
    "s1": (61, 2, 512, 6, 6, (512, 4, 3), True, True, 512, 4, 2, False),        # conv4_3 -> L2Norm -> fuse_11 -> loc/conf[0]
    "s1_train": (62, 2, 512, 7, 5, (512, 4, 3), True, True, 512, 4, 2, True),   # batch statistics
    "s2": (63, 1, 1024, 5, 5, (1024, 4, 1), True, False, 1024, 6, 2, False),    # conv7 (1x1, groups 4) -> fuse_21 -> loc/conf[1]
    "s4": (64, 2, 256, 5, 5, None, True, False, 256, 6, 2, False),              # extras source: fuse_41 -> loc/conf[3]
    "s1_nobn": (65, 1, 512, 4, 9, (512, 4, 3), False, True, 512, 4, 3, False),  # batch_norm=False variant, 3 classes
}


def block_case(tag):
    """-> (x[N,C,H,W] fp32 (bf16-representable, post-ReLU statistics), prm dict for oracle.source_block, training)"""
    from oracle.source_block import bf16_round
    seed, N, C, H, W, gc, bn, l2, Cf, A, ncls, training = BLOCK_CASES[tag]
    r = np.random.RandomState(seed)
    x = bf16_round(np.maximum(r.randn(N, C, H, W), 0).astype(np.float32))
    prm = {"bn_eps": 1e-5}

    def conv(co, ci, k, name):
        fan = ci * k * k
        prm[name + "_w"] = (r.randn(co, ci, k, k) * np.sqrt(2.0 / fan)).astype(np.float32)
        prm[name + "_b"] = (r.randn(co) * 0.1).astype(np.float32)

    def norm(c, name):
        prm[name + "_w"] = r.uniform(0.5, 1.5, c).astype(np.float32)
        prm[name + "_b"] = (r.randn(c) * 0.1).astype(np.float32)
        prm[name + "_mean"] = (r.randn(c) * 0.2).astype(np.float32)
        prm[name + "_var"] = r.uniform(0.5, 1.5, c).astype(np.float32)

    c_mid = C
    if gc is not None:
        co, groups, k = gc
        conv(co, C // groups, k, "gconv")
        prm["groups"], prm["gconv_pad"] = groups, (k - 1) // 2
        if bn:
            norm(co, "bn")
        c_mid = co
    if l2:
        prm["l2norm_w"] = (20.0 * r.uniform(0.8, 1.2, c_mid)).astype(np.float32)
    conv(Cf, c_mid, 1, "fuse")
    if bn:
        norm(Cf, "bn_fuse")
    conv(A * 4, Cf, 3, "loc")
    conv(A * ncls, Cf, 3, "conf")
    return x, prm, training


def block_upstream(tag, n_loc, n_conf):
    """seeded upstream gradients (d_loc[N, n_loc], d_conf[N, n_conf]) of a source-block case, float64"""
    seed, N = BLOCK_CASES[tag][0], BLOCK_CASES[tag][1]
    r = np.random.RandomState(seed + 1000)
    return r.randn(N, n_loc), r.randn(N, n_conf)


# ---- GSSD++'s modulated deformable convolution (layers/dcn_v2_custom.py) ----------------------------------------------------------
# tag -> (seed, N, C_in, C_out, H, W, deformable groups, scale of the offset convolution's weights; 0 = the module's own
# zero initialisation (dcn_v2_custom.py:72-74): offsets exactly 0, masks exactly 0.5, samples ON the pixel grid and on y = -1)
DCN_CASES = {
    "rand": (501, 2, 128, 64, 7, 7, 4, 0.02),
    "zero": (502, 1, 128, 64, 5, 6, 2, 0.0),
    "wide": (503, 1, 128, 64, 9, 5, 1, 0.05),          # large offsets: many samples leave the image
}


def dcn_case(tag):
    """-> dict(x, weight, bias, com_w, com_b, gout, dg): inputs, the parameters of a `DCN` module (dcn_v2_custom.py:58-88) and
    the upstream gradient of its output."""
    seed, N, C, O, H, W, dg, s = DCN_CASES[tag]
    r = np.random.RandomState(seed)
    f = lambda *shape: r.standard_normal(shape).astype(np.float32)
    return dict(x=f(N, C, H, W), weight=f(O, C, 3, 3) / np.float32(np.sqrt(9 * C)), bias=0.1 * f(O),
                com_w=np.float32(s) * f(dg * 27, C, 3, 3), com_b=(np.float32(s * 10) * f(dg * 27)), gout=f(N, O, H, W), dg=dg)


def strided_sample(a, stride=5):
    """what the DCN fixture keeps of a large gradient: every `stride`-th element, the sum and the absolute sum"""
    a = np.asarray(a, np.float64).ravel()
    return np.concatenate([a[::stride], [a.sum(), np.abs(a).sum()]])


# ---- GSSD++'s Self_Attn block (layers/self_attn.py) -----------------------------------------------------------------------------
# tag -> (seed, B, C, H, max_pool_factor)
SA_CASES = {
    "full": (601, 2, 256, 5, 1),          # keys = queries (the configuration GSSD++ trains with, max_pool_factor 1)
    "pooled": (602, 1, 256, 6, 2),        # 3 x 3 pooled keys against 36 queries
    "single": (603, 2, 256, 1, 4),        # the 1 x 1 map of the last source: one query, one key
}


def sa_case(tag):
    """-> (x, state dict of a Self_Attn module (the reference's names), upstream gradients of its first two outputs)"""
    seed, B, C, H, _ = SA_CASES[tag]
    r = np.random.RandomState(seed)
    f = lambda *shape: r.standard_normal(shape).astype(np.float32)
    prm = {"sigma": np.float32([0.7])}
    for name, co, ci in (("snconv1x1_theta", C // 8, C), ("snconv1x1_phi", C // 8, C), ("snconv1x1_g", C // 2, C), ("snconv1x1_attn", C, C // 2)):
        w, u = f(co, ci, 1, 1) / np.float32(np.sqrt(ci)), f(co)
        for _ in range(8):                                   # a few power iterations, as a trained module's buffers would hold
            v = w.reshape(co, ci).T @ u; v /= np.linalg.norm(v)
            u = w.reshape(co, ci) @ v; u /= np.linalg.norm(u)
        prm[name + ".weight_orig"], prm[name + ".bias"] = w, 0.1 * f(co)
        prm[name + ".weight_u"], prm[name + ".weight_v"] = u.astype(np.float32), v.astype(np.float32)
    return f(B, C, H, H), prm, (f(B, C, H, H), f(B, C, H, H))


# ---- the grouped backbone triples conv3_2 .. conv5_3 under autograd (source_block.PMConvLayer) ---------------------------------------
# tag -> (seed, first index into the reference's vgg list, number of [Conv2d, BatchNorm2d, ReLU] triples, N, H, W)
BACKBONE_CASES = {
    "conv3_2-3": (701, 17, 2, 2, 9, 11),        # 256 -> 256 -> 256, 64 channels per group
    "conv4_1-2": (702, 24, 2, 2, 8, 7),         # 256 -> 512 -> 512
    "conv5_1-3": (703, 34, 3, 3, 6, 5),         # 512 -> 512 three times
}
BACKBONE_CHANNELS = {17: 256, 20: 256, 24: 256, 27: 512, 34: 512, 37: 512, 40: 512}      # input channels of vgg[k]; outputs: 256 / 512


def backbone_case(tag):
    """-> (x[N,C,H,W] fp32 (bf16-representable, post-ReLU), list of per-triple dicts (w, b, gamma, beta; conv weights
    bf16-representable), upstream gradient of the last ReLU output)"""
    from oracle.source_block import bf16_round
    seed, first, n_triples, N, H, W = BACKBONE_CASES[tag]
    r = np.random.RandomState(seed)
    c_in = BACKBONE_CHANNELS[first]
    x = bf16_round(np.maximum(r.randn(N, c_in, H, W), 0).astype(np.float32))
    prm = []
    for t in range(n_triples):
        k = first + 3 * t
        c_in = BACKBONE_CHANNELS[k]
        c_out = 256 if k < 24 else 512
        fan = (c_in // 4) * 9
        prm.append(dict(w=bf16_round((r.randn(c_out, c_in // 4, 3, 3) * np.sqrt(2.0 / fan)).astype(np.float32)),
                        b=(r.randn(c_out) * 0.1).astype(np.float32),
                        gamma=r.uniform(0.5, 1.5, c_out).astype(np.float32), beta=(r.randn(c_out) * 0.2).astype(np.float32)))
    gout = r.randn(N, prm[-1]["w"].shape[0], H, W).astype(np.float32)
    return x, prm, gout
