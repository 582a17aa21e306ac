"""PriorBox — same constructor/forward as layers/functions/prior_box.py:5-172 of the reference."""
import hashlib
import json
import os

import numpy as np
import torch

from ... import _lib


class PriorBox(object):
    """Default boxes in center-offset form for every source feature map (prior_box.py:5-12).

    The boxes are generated on the GPU in float64 in the reference's operation order and rounded to
    float32 once, which reproduces the Python-float loop of prior_box.py:32-172 bit for bit."""

    def __init__(self, cfg):
        super(PriorBox, self).__init__()
        self.image_size = cfg['min_dim']
        self.num_priors = len(cfg['aspect_ratios'])
        self.variance = cfg['variance'] or [0.1]
        self.feature_maps = cfg['feature_maps']
        self.min_sizes = cfg['min_sizes']
        self.max_sizes = cfg['max_sizes']
        self.steps = cfg['steps']
        self.aspect_ratios = cfg['aspect_ratios']
        self.clip = cfg['clip']
        self.version = cfg['name']
        for v in self.variance:                       # prior_box.py:28-30
            if v <= 0:
                raise ValueError('Variances must be greater than 0')
        self._cfg = dict(cfg)

    def _cache_file(self):
        """GSSD_PRIOR_CACHE=<dir>: the boxes the kernel produced are serialised there (one .npy per configuration), so that a
        model object — the reference builds its priors inside SSD.__init__, ssd_multiphase_custom_group.py:48-49 — can be
        CONSTRUCTED on a box without a GPU (checkpoint surgery, the CPU import tests).  Nothing is computed on the CPU."""
        d = os.environ.get("GSSD_PRIOR_CACHE")
        if not d:
            return None
        key = json.dumps({k: self._cfg[k] for k in sorted(self._cfg)}, sort_keys=True, default=str)
        return os.path.join(d, "priors_%s_%s.npy" % (self.version, hashlib.sha1(key.encode()).hexdigest()[:16]))

    def forward(self, device=None):
        """-> FloatTensor[P,4] on the CPU like the reference (prior_box.py:168), or on `device`."""
        cache = self._cache_file()
        if not torch.cuda.is_available() and cache is not None and os.path.exists(cache):
            return torch.from_numpy(np.load(cache))
        lib = _lib.require_cuda()
        c = _lib.prior_cfg(self._cfg)
        n = lib.gssd_priorbox_count(c)
        _lib.check(n if n < 0 else 0, 'Variances must be greater than 0' if n == _lib.ERR_VALUE else "")
        dev = _lib.device_of()
        with torch.cuda.device(dev):
            out = torch.empty((n, 4), dtype=torch.float32, device=dev)
            _lib.check(lib.gssd_priorbox(c, out.data_ptr(), _lib.stream()), "gssd_priorbox")
        if cache is not None and not os.path.exists(cache):
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            np.save(cache, out.cpu().numpy())
        return out if device is not None and torch.device(device).type == "cuda" else out.cpu()
