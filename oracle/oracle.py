"""numpy-facing ctypes wrapper of oracle/libgssd_oracle.so (gssd_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgssd_oracle.so")

MAX_MAPS, MAX_AR = 8, 8
PRIOR_V2, PRIOR_V2_CUSTOM, PRIOR_LEGACY = 0, 1, 2


class PriorCfg(C.Structure):
    """Mirror of `gssd_prior_cfg` (include/gssd.h)."""
    _fields_ = [
        ("version", C.c_int32), ("n_maps", C.c_int32), ("clip", C.c_int32),
        ("feature_maps", C.c_int32 * MAX_MAPS), ("n_ar", C.c_int32 * MAX_MAPS),
        ("min_dim", C.c_double), ("steps", C.c_double * MAX_MAPS),
        ("min_sizes", C.c_double * MAX_MAPS), ("max_sizes", C.c_double * MAX_MAPS),
        ("aspect_ratios", (C.c_double * MAX_AR) * MAX_MAPS), ("variance", C.c_double * 2),
    ]


def prior_cfg(cfg):
    """cfg dict (data/config.py:19-157 layout) -> PriorCfg; version chosen as prior_box.py:35-141."""
    c = PriorCfg()
    name = cfg["name"]
    if name in ("v2", "v2_512"):
        c.version = PRIOR_V2
    elif name in ("v2_custom", "v2_custom_squareonly", "v2_custom_512"):
        c.version = PRIOR_V2_CUSTOM
    else:
        c.version = PRIOR_LEGACY
    c.n_maps = len(cfg["feature_maps"])
    c.clip = 1 if cfg["clip"] else 0
    c.min_dim = float(cfg["min_dim"])
    for k in range(c.n_maps):
        c.feature_maps[k] = int(cfg["feature_maps"][k])
        c.steps[k] = float(cfg["steps"][k])
        c.min_sizes[k] = float(cfg["min_sizes"][k])
        c.max_sizes[k] = float(cfg["max_sizes"][k])
        ars = cfg["aspect_ratios"][k]
        c.n_ar[k] = len(ars)
        for a, ar in enumerate(ars):
            c.aspect_ratios[k][a] = float(ar)
    var = cfg["variance"] or [0.1]
    c.variance[0] = float(var[0])
    c.variance[1] = float(var[1]) if len(var) > 1 else float(var[0])
    return c


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "gssd_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def set_threads(n):
    """OpenMP threads the oracle's image loops use from now on (torchrun exports OMP_NUM_THREADS=1, which would
    otherwise silently make the CPU arm single-threaded); returns the count in effect."""
    lib()
    omp = C.CDLL("libgomp.so.1")                      # the instance the oracle library is linked against
    if n and n > 0:
        omp.omp_set_num_threads(int(n))
    return int(omp.omp_get_max_threads())


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def priorbox(cfg):
    c = prior_cfg(cfg)
    n = lib().gssd_oracle_priorbox_count(C.byref(c))
    if n == -4:
        raise ValueError("Variances must be greater than 0")
    if n < 0:
        raise RuntimeError("oracle priorbox_count: %d" % n)
    out = np.empty((n, 4), np.float32)
    rc = lib().gssd_oracle_priorbox(C.byref(c), _p(out))
    assert rc == 0, rc
    return out


def _unary(fn, boxes):
    b = _f(boxes)
    out = np.empty_like(b)
    getattr(lib(), fn)(_p(b), C.c_int(b.shape[0]), _p(out))
    return out


def point_form(boxes):
    return _unary("gssd_oracle_point_form", boxes)


def center_size(boxes):
    return _unary("gssd_oracle_center_size", boxes)


def _pair(fn, a, b):
    a, b = _f(a), _f(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    getattr(lib(), fn)(_p(a), C.c_int(a.shape[0]), _p(b), C.c_int(b.shape[0]), _p(out))
    return out


def intersect(a, b):
    return _pair("gssd_oracle_intersect", a, b)


def jaccard(a, b):
    return _pair("gssd_oracle_jaccard", a, b)


def encode(matched, priors, variances):
    m, p = _f(matched), _f(priors)
    out = np.empty_like(m)
    lib().gssd_oracle_encode(_p(m), _p(p), C.c_int(m.shape[0]), C.c_float(variances[0]),
                             C.c_float(variances[1]), _p(out))
    return out


def decode(loc, priors, variances):
    l, p = _f(loc), _f(priors)
    out = np.empty_like(l)
    lib().gssd_oracle_decode(_p(l), _p(p), C.c_int(l.shape[0]), C.c_float(variances[0]),
                             C.c_float(variances[1]), _p(out))
    return out


def log_sum_exp(x):
    x = _f(x)
    out = np.empty((x.shape[0], 1), np.float32)
    lib().gssd_oracle_log_sum_exp(_p(x), C.c_int(x.shape[0]), C.c_int(x.shape[1]), _p(out))
    return out


def match(threshold, truths, priors, variances, labels):
    """box_utils.match for one image -> dict(loc_t, conf_t, best_truth_idx, best_truth_overlap)."""
    t, p, lab = _f(truths), _f(priors), _f(labels)
    P = p.shape[0]
    loc_t = np.empty((P, 4), np.float32)
    conf_t = np.empty((P,), np.int64)
    bti = np.empty((P,), np.int32)
    bto = np.empty((P,), np.float32)
    rc = lib().gssd_oracle_match(C.c_float(threshold), _p(t), _p(lab), C.c_int(t.shape[0]), _p(p),
                                 C.c_int(P), C.c_float(variances[0]), C.c_float(variances[1]),
                                 _p(loc_t), _p(conf_t, C.c_int64), _p(bti, C.c_int32), _p(bto))
    if rc == -5:
        raise IndexError("match: image without ground truth")
    assert rc == 0, rc
    return dict(loc_t=loc_t, conf_t=conf_t, best_truth_idx=bti, best_truth_overlap=bto)


def pack_targets(targets):
    """list of [n_i,5] arrays -> (gt[sum,5] float32, gt_off[B+1] int32)."""
    off = np.zeros(len(targets) + 1, np.int32)
    for i, t in enumerate(targets):
        off[i + 1] = off[i] + np.asarray(t).shape[0]
    gt = np.concatenate([_f(t).reshape(-1, 5) for t in targets], 0) if len(targets) else np.zeros((0, 5), np.float32)
    return np.ascontiguousarray(gt, np.float32), off


def multibox_loss(loc, conf, priors, targets, threshold=0.5, negpos_ratio=3, variances=(0.1, 0.2),
                  grads=True, extras=True, x_max=None, n_total=None):
    """MultiBoxLoss.forward (+ autograd) -> dict.  x_max / n_total: the batch-global max of conf and number
    of positives when (loc, conf, targets) is one rank's shard of a larger batch."""
    loc, conf, priors = _f(loc), _f(conf), _f(priors)
    B, P, Cn = conf.shape
    gt, off = pack_targets(targets)
    losses = np.zeros(2, np.float32)
    num_pos = np.zeros(B, np.int32)
    stats = np.zeros(2, np.float32)
    r = dict()
    if extras:
        r["loc_t"] = np.empty((B, P, 4), np.float32)
        r["conf_t"] = np.empty((B, P), np.int64)
        r["pos"] = np.empty((B, P), np.uint8)
        r["neg"] = np.empty((B, P), np.uint8)
        r["key"] = np.empty((B, P), np.float32)
        r["kth_gap"] = np.empty((B,), np.float32)
    if grads:
        r["grad_loc"] = np.empty((B, P, 4), np.float32)
        r["grad_conf"] = np.empty((B, P, Cn), np.float32)
    g = r.get
    rc = lib().gssd_oracle_multibox_loss(
        _p(loc), _p(conf), _p(priors), C.c_int(B), C.c_int(P), C.c_int(Cn), _p(gt), _p(off, C.c_int32),
        C.c_float(threshold), C.c_int(negpos_ratio), C.c_float(variances[0]), C.c_float(variances[1]),
        _p(losses), _p(num_pos, C.c_int32), _p(g("loc_t")), _p(g("conf_t"), C.c_int64),
        _p(g("pos"), C.c_uint8), _p(g("neg"), C.c_uint8), _p(g("key")), _p(g("kth_gap")),
        _p(g("grad_loc")), _p(g("grad_conf")),
        C.byref(C.c_float(x_max)) if x_max is not None else None,
        C.byref(C.c_int32(n_total)) if n_total is not None else None, _p(stats))
    if rc == -5:
        raise IndexError("multibox_loss: image without ground truth")
    assert rc == 0, rc
    r.update(loss_l=losses[0], loss_c=losses[1], num_pos=num_pos, local_x_max=stats[0], local_n=int(stats[1]))
    return r


def nms(boxes, scores, overlap=0.5, top_k=200):
    b, s = _f(boxes).reshape(-1, 4), _f(scores).reshape(-1)
    n = s.shape[0]
    keep = np.zeros((n,), np.int64)
    margin = np.zeros(2, np.float32)
    cnt = lib().gssd_oracle_nms(_p(b), _p(s), C.c_int(n), C.c_float(overlap), C.c_int(top_k),
                                _p(keep, C.c_int64), _p(margin))
    return keep, cnt, float(min(margin[0], margin[1]))


def detect(loc, conf, priors, num_classes, top_k, conf_thresh, nms_thresh, variances=(0.1, 0.2)):
    loc, conf, priors = _f(loc), _f(conf), _f(priors)
    B, P = loc.shape[0], priors.shape[0]
    out = np.empty((B, num_classes, top_k, 5), np.float32)
    count = np.empty((B, num_classes), np.int32)
    keep_idx = np.empty((B, num_classes, top_k), np.int32)
    margin = np.full((B, num_classes), np.inf, np.float32)
    cut_gap = np.full((B, num_classes), np.inf, np.float32)
    rc = lib().gssd_oracle_detect(_p(loc), _p(conf), _p(priors), C.c_int(B), C.c_int(P),
                                  C.c_int(num_classes), C.c_int(top_k), C.c_float(conf_thresh),
                                  C.c_float(nms_thresh), C.c_float(variances[0]), C.c_float(variances[1]),
                                  _p(out), _p(count, C.c_int32), _p(keep_idx, C.c_int32), _p(margin), _p(cut_gap))
    if rc == -4:
        raise ValueError("nms_threshold must be non negative.")
    assert rc == 0, rc
    return dict(out=out, count=count, keep_idx=keep_idx, margin=margin, cut_gap=cut_gap)


def l2norm(x, weight, eps=1e-10):
    x, w = _f(x), _f(weight)
    B, Cn = x.shape[:2]
    HW = int(np.prod(x.shape[2:]))
    y = np.empty_like(x)
    lib().gssd_oracle_l2norm(_p(x), _p(w), C.c_int(B), C.c_int(Cn), C.c_int(HW), C.c_float(eps), _p(y))
    return y
