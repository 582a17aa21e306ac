"""box_utils — same functions and signatures as layers/box_utils.py of the reference, computed by
libgssd_b200.so.  Tensors may live on the CPU or on a CUDA device; results come back on the device of
the first argument.  There is no CPU implementation: a CUDA device is required."""
import numpy as np
import torch

from .. import _lib


def _run(first, fn):
    dev = _lib.device_of(first)
    with torch.cuda.device(dev):
        out = fn(dev)
    return out if (not isinstance(first, torch.Tensor)) or first.is_cuda else out.cpu()


def point_form(boxes):
    """(cx, cy, w, h) -> (xmin, ymin, xmax, ymax)   [box_utils.py:4-13]"""
    lib = _lib.require_cuda()

    def go(dev):
        b = _lib.f32(boxes, dev)
        out = torch.empty_like(b)
        _lib.check(lib.gssd_point_form(b.data_ptr(), b.size(0), out.data_ptr(), _lib.stream()))
        return out
    return _run(boxes, go)


def center_size(boxes):
    """(xmin, ymin, xmax, ymax) -> (cx, cy, w, h)   [box_utils.py:16-25, documented intent]"""
    lib = _lib.require_cuda()

    def go(dev):
        b = _lib.f32(boxes, dev)
        out = torch.empty_like(b)
        _lib.check(lib.gssd_center_size(b.data_ptr(), b.size(0), out.data_ptr(), _lib.stream()))
        return out
    return _run(boxes, go)


def _pair(name, box_a, box_b):
    lib = _lib.require_cuda()

    def go(dev):
        a, b = _lib.f32(box_a, dev), _lib.f32(box_b, dev)
        out = torch.empty((a.size(0), b.size(0)), dtype=torch.float32, device=dev)
        _lib.check(getattr(lib, name)(a.data_ptr(), a.size(0), b.data_ptr(), b.size(0), out.data_ptr(), _lib.stream()))
        return out
    return _run(box_a, go)


def intersect(box_a, box_b):
    """[A,4], [B,4] point form -> intersection areas [A,B]   [box_utils.py:28-46]"""
    return _pair("gssd_intersect", box_a, box_b)


def jaccard(box_a, box_b):
    """IoU matrix [A,B]; union = area_a + area_b - inter   [box_utils.py:49-67]"""
    return _pair("gssd_jaccard", box_a, box_b)


class _PinnedRing(object):
    """Small ring of pinned host staging buffers for the per-step ground-truth upload.  A slot is reused
    only after the copy that read it has completed (CUDA event); under CUDA-graph capture a fresh buffer
    is allocated and kept alive instead, because the graph re-reads it at every replay."""

    def __init__(self, slots=8):
        self.bufs = [None] * slots
        self.views = [None] * slots
        self.events = [None] * slots
        self.i = 0
        self.keep = []

    def get(self, nbytes):
        """-> (pinned uint8 tensor, numpy uint8 view of it, slot)"""
        if torch.cuda.is_current_stream_capturing():
            buf = torch.empty((max(nbytes, 16),), dtype=torch.uint8).pin_memory()
            self.keep.append(buf)
            return buf, buf.numpy(), None
        k = self.i
        self.i = (self.i + 1) % len(self.bufs)
        if self.events[k] is not None:
            self.events[k].synchronize()
        if self.bufs[k] is None or self.bufs[k].numel() < nbytes:
            self.bufs[k] = torch.empty((max(2 * nbytes, 4096),), dtype=torch.uint8).pin_memory()
            self.views[k] = self.bufs[k].numpy()
        return self.bufs[k], self.views[k], k

    def sent(self, k):
        if k is not None:
            if self.events[k] is None:
                self.events[k] = torch.cuda.Event()
            self.events[k].record()


_ring = _PinnedRing()


class PackedTargets(object):
    """Device image of the reference's `targets` list (multibox_loss.py:67-69): gt[sum_G,5] float32 rows
    (xmin,ymin,xmax,ymax,label) and gt_off[B+1] int32 row offsets, plus the host-known sizes."""
    __slots__ = ("gt", "gt_off", "sum_g", "g_max", "batch")

    def __init__(self, gt, gt_off, sum_g, g_max, batch):
        self.gt, self.gt_off, self.sum_g, self.g_max, self.batch = gt, gt_off, sum_g, g_max, batch

    def __iter__(self):          # (gt, gt_off, sum_g, g_max) = pack_targets(...)
        return iter((self.gt, self.gt_off, self.sum_g, self.g_max))


def _offsets(lens):
    if min(lens) <= 0:
        raise IndexError("match: an image has no ground-truth box")      # the reference fails at box_utils.py:94
    offs = np.zeros(len(lens) + 1, np.int32)
    np.cumsum(lens, out=offs[1:])
    return offs


def pack_target_list(targets, dev):
    """list of [n_i,5] tensors (the DataLoader format, data_custom_v2.py:260-263) -> PackedTargets on `dev`.
    CPU targets travel in ONE pinned-buffer H2D copy (rows + offsets); CUDA targets are concatenated on the
    device and only the offsets are uploaded.  No host synchronisation either way."""
    if isinstance(targets, PackedTargets):
        return targets
    B = len(targets)
    lens = [int(t.shape[0]) if t.dim() == 2 else 0 for t in targets]
    offs = _offsets(lens)
    sum_g, g_max = int(offs[B]), max(lens)
    on_cpu = not any(t.is_cuda for t in targets)
    n_rows = sum_g * 5 if on_cpu else 0
    nbytes = 4 * (n_rows + B + 1)
    buf, view, slot = _ring.get(nbytes)
    if on_cpu:
        np.concatenate([t.detach().numpy() for t in targets], axis=0,
                       out=view[:4 * n_rows].view(np.float32).reshape(sum_g, 5), casting="unsafe")
    view[4 * n_rows:nbytes].view(np.int32)[:] = offs
    staged = buf[:nbytes].to(dev, non_blocking=True)
    _ring.sent(slot)
    gt_off = staged[4 * n_rows:].view(torch.int32)
    if on_cpu:
        gt = staged[:4 * n_rows].view(torch.float32).view(sum_g, 5)
    else:
        gt = torch.cat([t.detach().to(dev) for t in targets], 0).to(torch.float32).contiguous()
    return PackedTargets(gt, gt_off, sum_g, g_max, B)


def pack_targets(truths_list, labels_list, dev):
    """per-image (truths[G,4], labels[G]) pairs -> PackedTargets (the argument form of box_utils.match)."""
    rows = [torch.cat([t.detach().reshape(-1, 4).float(), l.detach().reshape(-1, 1).float().to(t.device)], 1)
            for t, l in zip(truths_list, labels_list)]
    return pack_target_list(rows, dev)


def match(threshold, truths, priors, variances, labels, loc_t, conf_t, idx):
    """Match priors with ground truth, encode, and write loc_t[idx], conf_t[idx] in place
    [box_utils.py:70-111]."""
    lib = _lib.require_cuda()
    dev = _lib.device_of(truths, priors, loc_t)
    with torch.cuda.device(dev):
        pri = _lib.f32(priors, dev)
        P = pri.size(0)
        gt, gt_off, sum_g, g_max = pack_targets([truths], [labels], dev)
        direct = loc_t.is_cuda and conf_t.is_cuda and loc_t.dtype == torch.float32 and conf_t.dtype == torch.int64 \
            and loc_t[idx].is_contiguous() and conf_t[idx].is_contiguous()
        lo = loc_t[idx] if direct else torch.empty((P, 4), dtype=torch.float32, device=dev)
        co = conf_t[idx] if direct else torch.empty((P,), dtype=torch.int64, device=dev)
        _lib.check(lib.gssd_match(pri.data_ptr(), P, gt.data_ptr(), gt_off.data_ptr(), 1, sum_g, g_max,
                                  float(threshold), float(variances[0]), float(variances[1]),
                                  lo.data_ptr(), co.data_ptr(), None, None, 0, _lib.stream()), "match")
        if not direct:
            loc_t[idx] = lo.to(loc_t.device)
            conf_t[idx] = co.to(conf_t.device)


def match_batch(threshold, targets, priors, variances, return_idx=False):
    """Batched form of the loop at multibox_loss.py:67-72: targets = list of [n_i,5] tensors.
    -> loc_t[B,P,4] float32, conf_t[B,P] int64 (, best_truth_idx[B,P] int32) on the GPU."""
    lib = _lib.require_cuda()
    dev = _lib.device_of(priors, *targets)
    with torch.cuda.device(dev):
        pri = _lib.f32(priors, dev)
        P, B = pri.size(0), len(targets)
        gt, gt_off, sum_g, g_max = pack_target_list(targets, dev)
        loc_t = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
        conf_t = torch.empty((B, P), dtype=torch.int64, device=dev)
        bti = torch.empty((B, P), dtype=torch.int32, device=dev) if return_idx else None
        _lib.check(lib.gssd_match(pri.data_ptr(), P, gt.data_ptr(), gt_off.data_ptr(), B, sum_g, g_max,
                                  float(threshold), float(variances[0]), float(variances[1]),
                                  loc_t.data_ptr(), conf_t.data_ptr(), _lib.ptr(bti), None, 0, _lib.stream()), "match")
    return (loc_t, conf_t, bti) if return_idx else (loc_t, conf_t)


def _codec(name, x, priors, variances):
    lib = _lib.require_cuda()

    def go(dev):
        a, p = _lib.f32(x, dev), _lib.f32(priors, dev)
        out = torch.empty_like(a)
        _lib.check(getattr(lib, name)(a.data_ptr(), p.data_ptr(), a.size(0), float(variances[0]),
                                      float(variances[1]), out.data_ptr(), _lib.stream()))
        return out
    return _run(x, go)


def encode(matched, priors, variances):
    """point-form matched boxes + center-form priors -> regression targets   [box_utils.py:114-135]"""
    return _codec("gssd_encode", matched, priors, variances)


def decode(loc, priors, variances):
    """loc predictions + center-form priors -> point-form boxes   [box_utils.py:139-157]"""
    return _codec("gssd_decode", loc, priors, variances)


def log_sum_exp(x):
    """log(sum(exp(x - max(x)), 1, keepdim=True)) + max(x), max over the WHOLE tensor
    [box_utils.py:160-168]"""
    lib = _lib.require_cuda()

    def go(dev):
        a = _lib.f32(x, dev)
        out = torch.empty((a.size(0), 1), dtype=torch.float32, device=dev)
        ws = torch.empty(4, dtype=torch.int32, device=dev)
        _lib.check(lib.gssd_log_sum_exp(a.data_ptr(), a.size(0), a.size(1), out.data_ptr(), ws.data_ptr(), 16,
                                        _lib.stream()))
        return out
    return _run(x, go)


def nms(boxes, scores, overlap=0.5, top_k=200):
    """Greedy NMS over the top_k highest scores -> (keep LongTensor[n] zero padded, count)
    [box_utils.py:174-238].  Like the reference it returns the bare `keep` for an empty input."""
    lib = _lib.require_cuda()
    n = int(scores.size(0))
    if boxes.numel() == 0:                              # box_utils.py:186-188
        return scores.new_zeros((n,), dtype=torch.long)
    dev = _lib.device_of(boxes, scores)
    with torch.cuda.device(dev):
        b, s = _lib.f32(boxes, dev), _lib.f32(scores, dev)
        keep = torch.empty((n,), dtype=torch.int64, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        _lib.check(lib.gssd_nms(b.data_ptr(), s.data_ptr(), n, float(overlap), int(top_k), keep.data_ptr(),
                                count.data_ptr(), None, 0, _lib.stream()), "nms")
        cnt = int(count.item())                         # the reference returns a Python int (box_utils.py:238)
    return (keep if scores.is_cuda else keep.cpu()), cnt
