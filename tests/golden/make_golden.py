"""Generate tests/golden/*.npz by running the UNMODIFIED reference (torch CPU) on seeded inputs.

Runs only in the build container (needs /root/reference; the GPU box has no copy):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every fixture stores the reference's outputs; large inputs are regenerated in the tests from the
seeds recorded here (grouped_ssd_pytorch_b200/synthetic.py, frozen numpy RandomState streams), small
hand-built inputs are stored alongside the outputs.  The reference has no tests or golden vectors of
its own, so these files are what pins the oracle.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/ssd_liverdet")
warnings.filterwarnings("ignore")

import torch  # noqa: E402

import data as ref_data  # noqa: E402  (reference)
from layers import Detect, MultiBoxLoss, PriorBox, box_utils as BU  # noqa: E402  (reference)
from layers.functions.detection import Detect as LegacyDetect  # noqa: E402
from layers.modules import L2Norm  # noqa: E402

from grouped_ssd_pytorch_b200 import config as our_cfg  # noqa: E402
from grouped_ssd_pytorch_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(4)
T = torch.from_numpy
VAR = [0.1, 0.2]


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def small_priors():
    """A 3-map prior set (P = 5*5*4 + 3*3*6 + 1*4 = 158) for edge cases."""
    cfg = dict(our_cfg.v2)
    cfg.update(feature_maps=[5, 3, 1], steps=[60, 100, 300], min_sizes=[60, 150, 240],
               max_sizes=[150, 240, 315], aspect_ratios=[[2], [2, 3], [2]])
    return cfg


# ------------------------------------------------------------------------------------------------
def gen_priors():
    out = {}
    for name in ("v2", "v2_512", "v2_custom", "v2_custom_512", "v2_custom_squareonly", "v1"):
        ref = PriorBox(getattr(ref_data, name)).forward().numpy()
        # our config dicts must describe the same boxes as the reference's
        assert our_cfg.ALL[name] == getattr(ref_data, name), name
        out[name] = ref
    out["small"] = PriorBox(small_priors()).forward().numpy()
    noclip = dict(small_priors()); noclip["clip"] = False
    out["small_noclip"] = PriorBox(noclip).forward().numpy()
    save("priors", **out)


def gen_box_utils():
    r = syn.rng(7)
    a = syn.targets(r, 1, 7, 7)[0][:, :4]
    pri = PriorBox(small_priors()).forward()
    b = BU.point_form(pri).numpy()
    loc = syn.loc(r, 1, pri.shape[0])[0]
    matched = a[r.randint(0, 7, size=pri.shape[0])]
    x = syn.conf_logits(r, 1, 300, 3)[0] * 3
    save("box_utils",
         a=a, priors=pri.numpy(), loc=loc, matched=matched, x=x,
         point_form=b,
         intersect=BU.intersect(T(a), T(b)).numpy(),
         jaccard=BU.jaccard(T(a), T(b)).numpy(),
         encode=BU.encode(T(matched), pri, VAR).numpy(),
         decode=BU.decode(T(loc), pri, VAR).numpy(),
         log_sum_exp=BU.log_sum_exp(T(x)).numpy())


def ref_match(truths, labels, priors, thr=0.5):
    P = priors.shape[0]
    loc_t = torch.zeros(1, P, 4)
    conf_t = torch.zeros(1, P, dtype=torch.long)
    BU.match(thr, T(truths), T(priors), VAR, T(labels), loc_t, conf_t, 0)
    # the reference does not return the matched indices: recompute them the way match() does
    ov = BU.jaccard(T(truths), BU.point_form(T(priors)))
    bpi = ov.max(1)[1]
    bto, bti = ov.max(0)
    bto.index_fill_(0, bpi, 2)
    for j in range(bpi.size(0)):
        bti[bpi[j]] = j
    return loc_t[0].numpy(), conf_t[0].numpy(), bti.numpy().astype(np.int32)


def gen_match():
    out = {}
    pri = PriorBox(ref_data.v2).forward().numpy()
    r = syn.rng(11)
    # (a) random, full v2 prior set, G = 1..5 — inputs regenerated from seed 11 in the tests
    tg = syn.targets(r, 3, 1, 5)
    for i, t in enumerate(tg):
        loc_t, conf_t, bti = ref_match(t[:, :4].copy(), t[:, 4].copy(), pri)
        out["rand%d_conf_t" % i] = conf_t.astype(np.int8)
        out["rand%d_bti" % i] = bti.astype(np.int8)
        if i == 0:
            out["rand0_loc_t"] = loc_t
        else:
            out["rand%d_loc_t_pos" % i] = loc_t[conf_t > 0]
    # (b) stress: v2_512 priors, G = 32
    pri512 = PriorBox(ref_data.v2_512).forward().numpy()
    t = syn.targets(syn.rng(12), 1, 32, 32)[0]
    t[:, 4] = (np.arange(32) % 3).astype(np.float32)          # labels 0,1,2 -> conf 1,2,3
    loc_t, conf_t, bti = ref_match(t[:, :4].copy(), t[:, 4].copy(), pri512)
    out.update(s512_targets=t, s512_conf_t=conf_t.astype(np.int8), s512_bti=bti.astype(np.int8),
               s512_loc_t_pos=loc_t[conf_t > 0])
    # (c) hand-built edge cases on the small prior set
    sp = PriorBox(small_priors()).forward().numpy()
    pf = BU.point_form(T(sp)).numpy()
    edge = {
        # two GT that share the same best prior (identical boxes): the later row wins
        "shared": np.array([[0.1, 0.1, 0.5, 0.5, 0], [0.1, 0.1, 0.5, 0.5, 1], [0.6, 0.6, 0.9, 0.9, 0]], np.float32),
        # a zero-area GT: IoU 0 with every prior, still force-matched to prior 0
        "zero_iou": np.array([[0.3, 0.3, 0.3, 0.3, 0], [0.2, 0.2, 0.6, 0.7, 0]], np.float32),
        # a GT equal to a prior box: IoU exactly 1 there
        "exact": np.concatenate([pf[40], [0]]).astype(np.float32)[None],
        # all GT identical to each other and zero area: every argmax is a tie
        "all_tie": np.array([[0.5, 0.5, 0.5, 0.5, 0]] * 3, np.float32),
    }
    # IoU exactly at the threshold: priors with power-of-two geometry, GT covering half of prior 0
    # (prior 3 IS that GT, so prior 0 is not force-matched and its 0.5 meets `< threshold` unaided)
    pw = np.array([[0.5, 0.5, 0.5, 0.5], [0.25, 0.25, 0.25, 0.25], [0.75, 0.75, 0.125, 0.125],
                   [0.375, 0.5, 0.25, 0.5]], np.float32)
    thr_gt = np.array([[0.25, 0.25, 0.5, 0.75, 0], [0.6875, 0.6875, 0.8125, 0.8125, 0]], np.float32)
    for k, t in edge.items():
        loc_t, conf_t, bti = ref_match(t[:, :4].copy(), t[:, 4].copy(), sp)
        out.update({"edge_%s_targets" % k: t, "edge_%s_conf_t" % k: conf_t.astype(np.int8),
                    "edge_%s_bti" % k: bti.astype(np.int8), "edge_%s_loc_t" % k: loc_t})
    with np.errstate(all="ignore"):
        loc_t, conf_t, bti = ref_match(thr_gt[:, :4].copy(), thr_gt[:, 4].copy(), pw)
    out.update(edge_thr_priors=pw, edge_thr_targets=thr_gt, edge_thr_conf_t=conf_t.astype(np.int8),
               edge_thr_bti=bti.astype(np.int8), edge_thr_loc_t=loc_t)
    out["small_priors"] = sp
    save("match", **out)


def ref_loss(loc, conf, priors, targets, num_classes, ratio=3, thr=0.5):
    loc = T(loc).clone().requires_grad_()
    conf = T(conf).clone().requires_grad_()
    crit = MultiBoxLoss(num_classes, thr, True, 0, True, ratio, 0.5, False, False)
    ll, lc = crit((loc, conf, T(priors)), [T(t) for t in targets])
    (ll + lc).backward()
    return ll.item(), lc.item(), loc.grad.numpy(), conf.grad.numpy()


def sparse(prefix, g, out):
    nz = np.flatnonzero(g.reshape(-1))
    out[prefix + "_idx"] = nz.astype(np.int32)
    out[prefix + "_val"] = g.reshape(-1)[nz]


def gen_loss():
    out = {}
    pri = PriorBox(ref_data.v2).forward().numpy()
    P = pri.shape[0]
    # (a) the BASELINE config-2 shape at B=4: seed 21
    r = syn.rng(21)
    tg = syn.targets(r, 4, 1, 5)
    loc, conf = syn.loc(r, 4, P), syn.conf_logits(r, 4, P, 2)
    ll, lc, gl, gc = ref_loss(loc, conf, pri, tg, 2)
    out.update(a_loss=np.array([ll, lc], np.float64))
    sparse("a_grad_loc", gl, out); sparse("a_grad_conf", gc, out)
    # (b) 3 classes, ratio 2, G up to 8: seed 22
    r = syn.rng(22)
    tg = syn.targets(r, 3, 1, 8)
    for t in tg:
        t[:, 4] = r.randint(0, 2, size=t.shape[0])
    loc, conf = syn.loc(r, 3, P), syn.conf_logits(r, 3, P, 3) * 2
    ll, lc, gl, gc = ref_loss(loc, conf, pri, tg, 3, ratio=2)
    out.update(b_loss=np.array([ll, lc], np.float64), b_labels=np.concatenate([t[:, 4] for t in tg]))
    sparse("b_grad_loc", gl, out); sparse("b_grad_conf", gc, out)
    # (c) num_neg clamp at P-1: small prior set, many GT, ratio 4 -> 4*num_pos > P-1: seed 23
    sp = PriorBox(small_priors()).forward().numpy()
    r = syn.rng(23)
    tg = syn.targets(r, 2, 60, 64)
    loc, conf = syn.loc(r, 2, sp.shape[0]), syn.conf_logits(r, 2, sp.shape[0], 2)
    ll, lc, gl, gc = ref_loss(loc, conf, sp, tg, 2, ratio=4)
    out.update(c_loss=np.array([ll, lc], np.float64), c_grad_loc=gl, c_grad_conf=gc)
    # (d) v2_512 stress shape, B=2, G up to 32: seed 24
    pri512 = PriorBox(ref_data.v2_512).forward().numpy()
    r = syn.rng(24)
    tg = syn.targets(r, 2, 1, 32)
    loc, conf = syn.loc(r, 2, pri512.shape[0]), syn.conf_logits(r, 2, pri512.shape[0], 2)
    ll, lc, gl, gc = ref_loss(loc, conf, pri512, tg, 2)
    out.update(d_loss=np.array([ll, lc], np.float64))
    sparse("d_grad_loc", gl, out); sparse("d_grad_conf", gc, out)
    save("loss", **out)


def gen_nms():
    out = {}
    r = syn.rng(31)
    # clustered boxes so that suppression happens
    centers = r.uniform(0.2, 0.8, size=(12, 2))
    c = centers[r.randint(0, 12, size=600)] + r.standard_normal((600, 2)) * 0.02
    wh = r.uniform(0.1, 0.2, size=(600, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    scores = r.uniform(0.01, 1.0, size=600).astype(np.float32)
    for tag, n, ov, k in (("a", 600, 0.45, 200), ("b", 150, 0.5, 200), ("c", 600, 0.3, 50), ("d", 1, 0.45, 200)):
        keep, count = BU.nms(T(boxes[:n]), T(scores[:n]), ov, k)
        out[tag + "_keep"] = keep.numpy()
        out[tag + "_count"] = np.array(count)
        out[tag + "_args"] = np.array([n, ov, k], np.float64)
    out["boxes"], out["scores"] = boxes, scores
    save("nms", **out)


def gen_detect():
    out = {}
    pri = PriorBox(ref_data.v2).forward()
    P = pri.shape[0]
    # (a) sparse-realistic scores, thr 0.2: seed 41;  (b) dense scores + clustered loc: seed 42
    # (c) in-model call: thr 0.01 (models/...group.py:384), 3 classes: seed 43
    for tag, seed, B, C, shift, sigma, thr in (("a", 41, 2, 2, -4.0, 0.5, 0.2), ("b", 42, 2, 2, 0.0, 0.05, 0.2),
                                               ("c", 43, 1, 3, -3.0, 0.2, 0.01)):
        r = syn.rng(seed)
        loc = syn.loc(r, B, P, sigma)
        conf = syn.detect_scores(r, B, P, C, shift)
        o = Detect.apply(C, 0, 200, thr, 0.45, T(loc), T(conf), pri)
        out[tag + "_out"] = o.numpy()
        out[tag + "_args"] = np.array([seed, B, C, shift, sigma, thr], np.float64)
        if tag == "a":   # the legacy instance-style Detect must agree (detection.py:13-62)
            o2 = LegacyDetect(C, 0, 200, thr, 0.45).forward(T(loc), T(conf), pri)
            assert torch.equal(o, o2)
    # (d) a class with no candidate at all
    r = syn.rng(44)
    loc = syn.loc(r, 1, P)
    conf = syn.detect_scores(r, 1, P, 2, -30.0)
    o = Detect.apply(2, 0, 200, 0.2, 0.45, T(loc), T(conf), pri)
    assert float(o.abs().sum()) == 0.0
    out["d_out"] = o.numpy()
    save("detect", **out)


def gen_l2norm():
    r = syn.rng(51)
    x = r.standard_normal((2, 64, 5, 7)).astype(np.float32)
    m = L2Norm(64, 20)
    with torch.no_grad():
        m.weight.copy_(T(r.uniform(10, 30, size=64).astype(np.float32)))
    xt = T(x).clone().requires_grad_()
    y = m(xt)
    gy = r.standard_normal(y.shape).astype(np.float32)
    y.backward(T(gy))
    save("l2norm", x=x, weight=m.weight.detach().numpy(), y=y.detach().numpy(), gy=gy,
         gx=xt.grad.numpy(), gw=m.weight.grad.numpy())


if __name__ == "__main__":
    gen_priors()
    gen_box_utils()
    gen_match()
    gen_loss()
    gen_nms()
    gen_detect()
    gen_l2norm()
