"""MultiBoxLoss — same constructor and forward as layers/modules/multibox_loss.py:8-120 of the
reference; the whole forward (matching, encoding, smooth-L1, hard-negative mining, cross-entropy,
normalisation) and its backward run in two kernel launches of libgssd_b200.so."""
import torch
import torch.nn as nn

from ... import _lib
from ... import dist as gdist
from ...config import v2 as cfg
from ..box_utils import pack_target_list


class _MultiBoxLossFn(torch.autograd.Function):
    """(loc, conf) -> (loss_l, loss_c) with d(loss_l)/d(loc) and d(loss_c)/d(conf) produced by the
    forward kernel; backward only rescales them by the upstream gradients (a no-op kernel when both
    are 1, i.e. `(loss_l + loss_c).backward()`, train_lesion_multiphase_v2.py:247-248)."""

    @staticmethod
    def forward(ctx, loc, conf, priors, gt, gt_off, sum_g, g_max, threshold, negpos_ratio, variance, masks, group, owner):
        import ctypes
        lib = _lib.require_cuda()
        dev = loc.device
        B, P, C = conf.shape
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        st = _lib.stream()
        # DataParallel semantics of the reference: x_max and N are over the GLOBAL batch
        # (train_lesion_multiphase_v2.py:242-246 -> multibox_loss.py:117, box_utils.py:167): the statistics of every rank
        # travel either through the peer exchange (NVLink peer stores issued by the kernels) or through an all-gather
        # between the two stages.
        _, world, _ = gdist.world(group)
        ex = gdist.peer_exchange(group, owner=owner) if world > 1 else None
        losses = torch.empty((2,), dtype=torch.float32, device=dev)
        grad_loc = torch.empty_like(loc) if need_grad else None
        grad_conf = torch.empty_like(conf) if need_grad else None
        pos = torch.empty((B, P), dtype=torch.uint8, device=dev) if masks else None
        neg = torch.empty((B, P), dtype=torch.uint8, device=dev) if masks else None
        ws_bytes = lib.gssd_workspace_bytes(_lib.WS_LOSS, B, P, C, sum_g, 0)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        var0, var1 = float(variance[0]), float(variance[1])
        one_launch = (world == 1 or ex is not None) and lib.gssd_mbox_fused_supported(B, P, C, g_max) != 0
        if one_launch:
            # the whole forward + backward as ONE launch (csrc/fused.cu): every CTA of the batch is resident at once
            num_pos = torch.empty((B,), dtype=torch.int32, device=dev)
            state = _lib.fused_state(dev, st)
            _lib.check(lib.gssd_mbox_loss_fused(loc.data_ptr(), conf.data_ptr(), priors.data_ptr(), B, P, C,
                                                gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, float(threshold),
                                                int(negpos_ratio), var0, var1, state.data_ptr(),
                                                ctypes.byref(ex.x) if ex is not None else None,
                                                losses.data_ptr(), _lib.ptr(grad_loc), _lib.ptr(grad_conf),
                                                _lib.ptr(pos), _lib.ptr(neg), num_pos.data_ptr(), ws.data_ptr(), ws_bytes, st),
                       "gssd_mbox_loss_fused")
        else:
            tags = torch.empty((B, P), dtype=torch.int16, device=dev)
            stats = torch.empty((_lib.STATS_HEADER_BYTES + 4 * B,), dtype=torch.uint8, device=dev)
            gstats, n_g = None, 0
            if ex is not None:
                _lib.check(lib.gssd_mbox_match_x(priors.data_ptr(), P, conf.data_ptr(), C, gt.data_ptr(), gt_off.data_ptr(),
                                                 B, sum_g, g_max, float(threshold), tags.data_ptr(), stats.data_ptr(),
                                                 ctypes.byref(ex.x), st), "gssd_mbox_match_x")
            else:
                _lib.check(lib.gssd_mbox_match(priors.data_ptr(), P, conf.data_ptr(), C, gt.data_ptr(), gt_off.data_ptr(),
                                               B, sum_g, g_max, float(threshold), tags.data_ptr(), stats.data_ptr(), st),
                           "gssd_mbox_match")
            if world > 1 and ex is None:
                gstats = gdist.all_gather_headers(stats[:_lib.STATS_HEADER_BYTES], group)
                n_g = world
            if ex is not None:
                _lib.check(lib.gssd_mbox_loss_x(loc.data_ptr(), conf.data_ptr(), priors.data_ptr(), B, P, C,
                                                gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, tags.data_ptr(), stats.data_ptr(),
                                                ctypes.byref(ex.x), int(negpos_ratio), var0, var1,
                                                losses.data_ptr(), _lib.ptr(grad_loc), _lib.ptr(grad_conf),
                                                _lib.ptr(pos), _lib.ptr(neg), ws.data_ptr(), ws_bytes, st), "gssd_mbox_loss_x")
            else:
                _lib.check(lib.gssd_mbox_loss(loc.data_ptr(), conf.data_ptr(), priors.data_ptr(), B, P, C,
                                              gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, tags.data_ptr(), stats.data_ptr(),
                                              _lib.ptr(gstats), n_g, int(negpos_ratio), var0, var1,
                                              losses.data_ptr(), _lib.ptr(grad_loc), _lib.ptr(grad_conf),
                                              _lib.ptr(pos), _lib.ptr(neg), ws.data_ptr(), ws_bytes, st), "gssd_mbox_loss")
            num_pos = stats[_lib.STATS_HEADER_BYTES:].view(torch.int32)
        ctx.grads = (grad_loc, grad_conf)
        ctx.set_materialize_grads(False)       # no zero tensors for the outputs nobody differentiates (num_pos, an unused loss)
        aux = [t for t in (pos, neg, num_pos) if t is not None]
        ctx.mark_non_differentiable(*aux)
        return losses[0], losses[1], pos, neg, num_pos

    @staticmethod
    def backward(ctx, g_l, g_c, *_unused):
        """d(loss_l)/d(loc) and d(loss_c)/d(conf) were produced by the forward kernel; they are scaled by the upstream gradients.
        `(loss_l + loss_c).backward()` (train_lesion_multiphase_v2.py:247-248; both upstream gradients present) scales the
        buffers in place — a kernel that returns at once when both are 1 — and hands them to autograd without a copy.  When only
        one of the two losses is being differentiated (`loss_l.backward(retain_graph=True); loss_c.backward()`,
        `torch.autograd.grad(loss_c, conf)`) the buffers are kept intact and a scaled copy of the one that takes part is returned,
        so that the other loss can still be differentiated afterwards."""
        if ctx.grads is None:
            raise RuntimeError("MultiBoxLoss: the gradient buffers of this forward were already handed to autograd by a backward "
                               "through both losses; differentiate loss_l and loss_c separately, or call forward again")
        grad_loc, grad_conf = ctx.grads
        if grad_loc is None:
            return (None,) * 13
        lib = _lib.load()
        dev = grad_loc.device

        def scalar(g):
            return g if (g.is_cuda and g.dtype == torch.float32) else g.to(device=dev, dtype=torch.float32)

        with torch.cuda.device(dev):
            if g_l is None or g_c is None:     # one loss only: out of place, the buffers stay valid for the other one
                out_l = grad_loc * scalar(g_l) if g_l is not None else None
                out_c = grad_conf * scalar(g_c) if g_c is not None else None
                return (out_l, out_c) + (None,) * 11
            ctx.grads = None                   # hand the buffers over: autograd can adopt them without a copy
            g_l, g_c = scalar(g_l), scalar(g_c)
            _lib.check(lib.gssd_mbox_scale_grads(grad_loc.data_ptr(), grad_loc.numel(), grad_conf.data_ptr(),
                                                 grad_conf.numel(), g_l.data_ptr(), g_c.data_ptr(), _lib.stream()),
                       "gssd_mbox_scale_grads")
        return (grad_loc, grad_conf) + (None,) * 11


class MultiBoxLoss(nn.Module):
    """SSD weighted loss (multibox_loss.py:8-29): L = (Lconf + Lloc) / N with
    1) prior <-> ground-truth matching at `overlap_thresh`, 2) variance-encoded offsets,
    3) hard-negative mining at `neg_pos`:1.

    Constructor arguments are the reference's (multibox_loss.py:31-44); only `num_classes`,
    `overlap_thresh` and `neg_pos` influence the result there, and the same holds here.

    Extra attributes (not in the reference):
      process_group : None = default group when torch.distributed is initialised with world_size > 1
                      (x_max and N then span the global batch, as under the reference's DataParallel);
                      False = keep both statistics local to this rank.
      keep_masks    : also produce the positive / hard-negative masks (`last_masks`) for inspection.
    With world_size > 1 the returned losses hold this rank's numerators over the global N: their sum
    over ranks is the reference's loss."""

    def __init__(self, num_classes, overlap_thresh, prior_for_matching,
                 bkg_label, neg_mining, neg_pos, neg_overlap, encode_target,
                 use_gpu=True):
        super(MultiBoxLoss, self).__init__()
        self.use_gpu = use_gpu
        self.num_classes = num_classes
        self.threshold = overlap_thresh
        self.background_label = bkg_label
        self.encode_target = encode_target
        self.use_prior_for_matching = prior_for_matching
        self.do_neg_mining = neg_mining
        self.negpos_ratio = neg_pos
        self.neg_overlap = neg_overlap
        self.variance = cfg['variance']          # multibox_loss.py:44
        self.process_group = None
        self.keep_masks = False
        self.last_masks = None
        self._prior_cache = {}

    def _device_priors(self, priors, dev):
        if priors.is_cuda and priors.dtype == torch.float32 and priors.is_contiguous() and priors.device == dev:
            return priors.detach()
        key = (priors.data_ptr(), tuple(priors.shape), priors._version, str(dev))
        hit = self._prior_cache.get(key)
        if hit is None:
            self._prior_cache.clear()
            hit = _lib.f32(priors, dev)
            self._prior_cache[key] = hit
        return hit

    def forward(self, predictions, targets):
        """predictions = (loc[B,P,4], conf[B,P,C], priors[P,4]); targets = list of [n_i,5] tensors
        (xmin,ymin,xmax,ymax,label), CPU or CUDA (multibox_loss.py:46-57) -> (loss_l, loss_c)."""
        _lib.require_cuda()
        loc_data, conf_data, priors = predictions
        num = loc_data.size(0)
        priors = priors[:loc_data.size(1), :]                  # multibox_loss.py:60
        dev = _lib.device_of(loc_data, conf_data)
        out_dev = loc_data.device
        with torch.cuda.device(dev):
            loc = loc_data.to(device=dev, dtype=torch.float32).contiguous()
            conf = conf_data.to(device=dev, dtype=torch.float32).contiguous()
            if conf.dim() == 2:
                conf = conf.view(num, -1, self.num_classes)
            pri = self._device_priors(priors, dev)
            gt, gt_off, sum_g, g_max = pack_target_list(
                targets if not isinstance(targets, (list, tuple)) else [targets[i] for i in range(num)], dev)
            loss_l, loss_c, pos, neg, num_pos = _MultiBoxLossFn.apply(
                loc, conf, pri, gt, gt_off, sum_g, g_max, self.threshold, self.negpos_ratio, self.variance,
                self.keep_masks, self.process_group, self)
            if self.keep_masks:
                self.last_masks = dict(pos=pos, neg=neg, num_pos=num_pos)
        if out_dev != dev:
            loss_l, loss_c = loss_l.to(out_dev), loss_c.to(out_dev)
        return loss_l, loss_c
