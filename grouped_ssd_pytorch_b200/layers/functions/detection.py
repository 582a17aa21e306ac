"""Detect — the final layer of SSD at test time: decode, per-class confidence threshold, top_k, NMS.

Serves both call styles of the reference:
  new style   Detect.apply(num_classes, bkg_label, top_k, conf_thresh, nms_thresh, loc, conf, priors)
              (layers/functions/detection_pytorch_ver_1point5.py:33; models/...custom_group.py:75,384)
  legacy      Detect(num_classes, bkg_label, top_k, conf_thresh, nms_thresh)(loc, conf, priors)
              (layers/functions/detection.py:13-24)
One kernel launch covers the whole batch (libgssd_b200.so: gssd_detect)."""
import torch

from ... import _lib
from ...config import v2 as cfg


def _detect(num_classes, top_k, conf_thresh, nms_thresh, loc_data, conf_data, prior_data, variance,
            want_aux=False, logits=False, class_bias=None):
    if nms_thresh <= 0:                                   # detection_pytorch_ver_1point5.py:39-40
        raise ValueError('nms_threshold must be non negative.')
    lib = _lib.require_cuda()
    dev = _lib.device_of(loc_data, conf_data, prior_data)
    num = loc_data.size(0)
    num_priors = prior_data.size(0)
    with torch.cuda.device(dev):
        loc = _lib.f32(loc_data, dev).view(num, num_priors, 4)
        conf = _lib.f32(conf_data, dev).view(num, num_priors, num_classes)   # ...1point5.py:58-59
        pri = _lib.f32(prior_data, dev)
        out = torch.empty((num, num_classes, top_k, 5), dtype=torch.float32, device=dev)
        count = torch.empty((num, num_classes), dtype=torch.int32, device=dev) if want_aux else None
        keep = torch.empty((num, num_classes, top_k), dtype=torch.int32, device=dev) if want_aux else None
        if logits:
            import ctypes
            bias = None
            if class_bias is not None:
                if len(class_bias) != num_classes:
                    raise ValueError("class_bias needs one entry per class")
                bias = (ctypes.c_float * num_classes)(*[float(v) for v in class_bias])
            _lib.check(lib.gssd_detect_logits(loc.data_ptr(), conf.data_ptr(), bias, pri.data_ptr(), num, num_priors,
                                              num_classes, int(top_k), float(conf_thresh), float(nms_thresh),
                                              float(variance[0]), float(variance[1]), out.data_ptr(), _lib.ptr(count),
                                              _lib.ptr(keep), _lib.stream()), "gssd_detect_logits")
        else:
            _lib.check(lib.gssd_detect(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), num, num_priors, num_classes,
                                       int(top_k), float(conf_thresh), float(nms_thresh), float(variance[0]),
                                       float(variance[1]), out.data_ptr(), _lib.ptr(count), _lib.ptr(keep),
                                       _lib.stream()), "gssd_detect")
    if not loc_data.is_cuda:
        out = out.cpu()
    return (out, count, keep) if want_aux else out


class Detect(object):
    """Output: FloatTensor[batch, num_classes, top_k, 5] rows (score, xmin, ymin, xmax, ymax) in
    descending score, zero padded; class 0 (background) is all zero."""

    def __init__(self, num_classes=None, bkg_label=0, top_k=200, conf_thresh=0.01, nms_thresh=0.45):
        self.num_classes = num_classes
        self.background_label = bkg_label
        self.top_k = top_k
        self.nms_thresh = nms_thresh
        if nms_thresh <= 0:                               # detection.py:19-20
            raise ValueError('nms_threshold must be non negative.')
        self.conf_thresh = conf_thresh
        self.variance = cfg['variance']

    # legacy instance style -------------------------------------------------------------------------
    def forward(self, loc_data, conf_data, prior_data):
        if self.num_classes is None:
            raise TypeError("Detect(): construct with (num_classes, bkg_label, top_k, conf_thresh, nms_thresh) "
                            "or use Detect.apply(...)")
        with torch.no_grad():
            return _detect(self.num_classes, self.top_k, self.conf_thresh, self.nms_thresh,
                           loc_data, conf_data, prior_data, self.variance)

    __call__ = forward

    # new style ---------------------------------------------------------------------------------------
    @staticmethod
    def apply(num_classes, bkg_label, top_k, conf_thresh, nms_thresh, loc_data, conf_data, prior_data):
        with torch.no_grad():
            return _detect(num_classes, top_k, conf_thresh, nms_thresh, loc_data, conf_data, prior_data,
                           cfg['variance'])

    @staticmethod
    def apply_logits(num_classes, bkg_label, top_k, conf_thresh, nms_thresh, loc_data, conf_logits, prior_data, class_bias=None):
        """`Detect.apply(..., loc, softmax(conf_logits + class_bias), priors)` in one kernel: the softmax of the model's test
        phase (ssd_multiphase_custom_group.py:384-390) is evaluated inside the threshold pass (not in the reference API)."""
        with torch.no_grad():
            return _detect(num_classes, top_k, conf_thresh, nms_thresh, loc_data, conf_logits, prior_data,
                           cfg['variance'], logits=True, class_bias=class_bias)

    @staticmethod
    def apply_logits_with_indices(num_classes, bkg_label, top_k, conf_thresh, nms_thresh, loc_data, conf_logits, prior_data,
                                  class_bias=None):
        """`apply_logits` + (count[B,C], keep_idx[B,C,top_k]) as `apply_with_indices`."""
        with torch.no_grad():
            return _detect(num_classes, top_k, conf_thresh, nms_thresh, loc_data, conf_logits, prior_data,
                           cfg['variance'], want_aux=True, logits=True, class_bias=class_bias)

    @staticmethod
    def apply_with_indices(num_classes, bkg_label, top_k, conf_thresh, nms_thresh, loc_data, conf_data, prior_data):
        """-> (output, count[B,C] int32, keep_idx[B,C,top_k] int32 prior indices, -1 padded)."""
        with torch.no_grad():
            return _detect(num_classes, top_k, conf_thresh, nms_thresh, loc_data, conf_data, prior_data,
                           cfg['variance'], want_aux=True)
