"""HostPipeline — the training-step call of the hot path with HOST buffers, `depth` steps in flight
(include/gssd.h: gssd_pipe_*).  One native call per step: the H2D of train_lesion_multiphase_v2.py:198-200, the criterion
of :246 (MultiBoxLoss forward + the gradients of its two outputs) and Detect (ssd_multiphase_custom_group.py:384-390).

    pipe = HostPipeline(B, priors, num_classes=2, depth=3)
    bufs = [pipe.host_buffers() for _ in range(3)]            # pinned: loc / conf / scores / losses / detections
    t = pipe.submit(bufs[i], targets)                           # targets: list of CPU [n_i,5] tensors
    pipe.wait(t);  bufs[i].losses, bufs[i].detections           # results on the host
    g_loc, g_conf = pipe.grads(t)                               # device tensors, valid until the slot is reused

With torch.distributed initialised (world_size > 1) the 16-byte loss statistics of every rank are all-gathered between
the two halves of a step, so that N and the LSE max span the global batch as under the reference's DataParallel.
No CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import dist as gdist


class HostBuffers(object):
    """page-locked host memory of one step in the layout of a pipeline slot — loc | conf | gt rows | row offsets (| scores) — so that
    a step's inputs travel in ONE H2D transfer; losses[2], detections[B,C,top_k,5]"""

    def __init__(self, B, P, Cn, top_k, with_scores=True, max_gt_rows=None):
        a = lambda n: (n + 255) // 256 * 256
        n_loc, n_conf = B * P * 4 * 4, B * P * Cn * 4
        self.max_gt_rows = int(max_gt_rows or B * _lib.MAX_GT_PER_IMAGE)
        n_gt, n_off = self.max_gt_rows * 5 * 4, (B + 1) * 4
        o_conf, o_gt = a(n_loc), a(n_loc) + a(n_conf)
        o_off = o_gt + a(n_gt)
        o_sc = o_off + a(n_off)
        self.arena = torch.empty((o_sc + (n_conf if with_scores else 0),), dtype=torch.uint8).pin_memory()
        self.loc = self.arena[:n_loc].view(torch.float32).view(B, P, 4)
        self.conf = self.arena[o_conf:o_conf + n_conf].view(torch.float32).view(B, P, Cn)
        self.gt = self.arena[o_gt:o_gt + n_gt].view(torch.float32)
        self.gt_off = self.arena[o_off:o_off + n_off].view(torch.int32)
        self.scores = self.arena[o_sc:].view(torch.float32).view(B, P, Cn) if with_scores else None
        self.losses = torch.zeros((2,), dtype=torch.float32).pin_memory()
        self.detections = torch.zeros((B, Cn, top_k, 5), dtype=torch.float32).pin_memory()
        self.gt_np, self.gt_off_np = self.gt.numpy(), self.gt_off.numpy()


class HostPipeline(object):
    def __init__(self, batch, priors, num_classes=2, top_k=200, depth=3, match_thresh=0.5, negpos_ratio=3,
                 variance=(0.1, 0.2), conf_thresh=0.01, nms_thresh=0.45, max_gt_rows=None, process_group=None, device=None,
                 detect_logits=False, class_bias=None):
        if nms_thresh <= 0:
            raise ValueError('nms_threshold must be non negative.')
        lib = _lib.require_cuda()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.priors = _lib.f32(priors, self.dev)
        self.B, self.P, self.C, self.top_k, self.depth = int(batch), int(self.priors.shape[0]), int(num_classes), int(top_k), int(depth)
        cfg = _lib.PipeCfg()
        cfg.B, cfg.P, cfg.C, cfg.top_k, cfg.depth = self.B, self.P, self.C, self.top_k, self.depth
        cfg.max_gt_rows = int(max_gt_rows or self.B * 32)
        cfg.match_thresh, cfg.var0, cfg.var1 = float(match_thresh), float(variance[0]), float(variance[1])
        cfg.conf_thresh, cfg.nms_thresh, cfg.negpos_ratio = float(conf_thresh), float(nms_thresh), int(negpos_ratio)
        self.cfg = cfg
        with torch.cuda.device(self.dev):
            nbytes = lib.gssd_pipe_arena_bytes(C.byref(cfg))
            if nbytes == 0:
                raise ValueError("HostPipeline: bad configuration")
            self.arena = torch.empty((nbytes,), dtype=torch.uint8, device=self.dev)
            h = C.c_void_p()
            _lib.check(lib.gssd_pipe_create(C.byref(h), C.byref(cfg), self.priors.data_ptr(), self.arena.data_ptr(), nbytes), "gssd_pipe_create")
        self._h, self._lib = h, lib
        self.detect_logits = bool(detect_logits)
        if self.detect_logits:                                    # Detect = softmax(conf + class_bias) fused: no scores upload
            bias = None
            if class_bias is not None:
                bias = (C.c_float * self.C)(*[float(v) for v in class_bias])
            _lib.check(lib.gssd_pipe_set_detect_logits(h, 1, bias), "gssd_pipe_set_detect_logits")
        self.group = process_group
        self._gathered = None
        _, world, _ = gdist.world(process_group)
        self._ex = gdist.peer_exchange(process_group, owner=self) if world > 1 else None
        if self._ex is not None:                                  # statistics over NVLink peer memory: one native call per step
            _lib.check(lib.gssd_pipe_set_xchg(h, C.byref(self._ex.x)), "gssd_pipe_set_xchg")

    def host_buffers(self):
        return HostBuffers(self.B, self.P, self.C, self.top_k, with_scores=not self.detect_logits, max_gt_rows=self.cfg.max_gt_rows)

    def _pack(self, bufs, targets):
        B = self.B
        if len(targets) != B:
            raise ValueError("expected %d target tensors, got %d" % (B, len(targets)))
        lens = [int(t.shape[0]) for t in targets]
        if min(lens) <= 0:
            raise IndexError("match: an image has no ground-truth box")      # the reference fails at box_utils.py:94
        sum_g, g_max = sum(lens), max(lens)
        if sum_g > self.cfg.max_gt_rows or g_max > _lib.MAX_GT_PER_IMAGE:
            raise RuntimeError("HostPipeline: more ground-truth rows than max_gt_rows")
        gt = bufs.gt_np[:sum_g * 5].reshape(sum_g, 5)
        np.concatenate([t.numpy() for t in targets], axis=0, out=gt, casting="unsafe")
        off = bufs.gt_off_np
        off[0] = 0
        np.cumsum(lens, out=off[1:])
        return bufs.gt.data_ptr(), bufs.gt_off.data_ptr(), sum_g, g_max

    def submit(self, bufs, targets, detect=True):
        """enqueue one step on `bufs` (its loc/conf/scores are read, its losses/detections written); returns the ticket.
        targets=None: a Detect-only step (inference) — no matching, no loss."""
        lib, h = self._lib, self._h
        if targets is None:
            if not detect:
                raise ValueError("HostPipeline.submit: nothing to do (no targets and detect=False)")
            sc = bufs.scores.data_ptr() if bufs.scores is not None else None
            t = lib.gssd_pipe_submit(h, bufs.loc.data_ptr(), bufs.conf.data_ptr(), sc, None, None, 0, 0, None,
                                     bufs.detections.data_ptr())
            if t < 0:
                self._raise(t, "gssd_pipe_submit")
            return t
        gt_p, off_p, sum_g, g_max = self._pack(bufs, targets)
        sc = bufs.scores.data_ptr() if (detect and bufs.scores is not None) else None
        det = bufs.detections.data_ptr() if detect else None
        _, world, _ = gdist.world(self.group)
        if world <= 1 or self._ex is not None:
            t = lib.gssd_pipe_submit(h, bufs.loc.data_ptr(), bufs.conf.data_ptr(), sc, gt_p, off_p, sum_g, g_max,
                                     bufs.losses.data_ptr(), det)
            if t < 0:
                self._raise(t, "gssd_pipe_submit")
            return t
        st = C.c_void_p()
        t = lib.gssd_pipe_begin(h, bufs.loc.data_ptr(), bufs.conf.data_ptr(), sc, gt_p, off_p, sum_g, g_max, det, C.byref(st))
        if t < 0:
            self._raise(t, "gssd_pipe_begin")
        slot = self.slot(t % self.depth)
        hdr = self._view(slot.stats, _lib.STATS_HEADER_BYTES, torch.uint8)
        with torch.cuda.stream(torch.cuda.ExternalStream(st.value, device=self.dev)):
            g = gdist.all_gather_headers(hdr, self.group)
        self._gathered = g                                        # keep alive until the next step
        _lib.check(lib.gssd_pipe_finish(h, t, g.data_ptr(), world, bufs.losses.data_ptr()), "gssd_pipe_finish")
        return t

    def wait(self, ticket):
        _lib.check(self._lib.gssd_pipe_wait(self._h, ticket), "gssd_pipe_wait")

    def slot(self, k):
        s = _lib.PipeSlot()
        _lib.check(self._lib.gssd_pipe_slot_info(self._h, int(k), C.byref(s)), "gssd_pipe_slot_info")
        return s

    def _view(self, ptr, nbytes, dtype):
        off = ptr - self.arena.data_ptr()
        return self.arena[off:off + nbytes].view(dtype)

    def grads(self, ticket):
        """(grad_loc[B,P,4], grad_conf[B,P,C]) of that step: views into the pipeline's arena, valid until the slot is reused"""
        s = self.slot(ticket % self.depth)
        BP = self.B * self.P
        return (self._view(s.grad_loc, BP * 16, torch.float32).view(self.B, self.P, 4),
                self._view(s.grad_conf, BP * self.C * 4, torch.float32).view(self.B, self.P, self.C))

    def _raise(self, code, what):
        code = int(code)
        if code <= -1000:
            code = -code - 1000
        _lib.check(code, what)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.gssd_pipe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
