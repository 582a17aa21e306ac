"""STAGED for the next round, not collected by pytest: the data gradient of the source block's grouped 3x3 and
dense 1x1 convolutions on the EXISTING forward kernel — `gssd_conv_igemm` on dY with `source_block.dgrad_weight(w)` — against the
numpy oracle (`oracle.source_block.conv2d_backward`).  Has not run yet.

    gpurun --timeout 200 -- 'timeout 120 python tools/next_round/check_dgrad.py'

Tolerance: operands are rounded to bf16 on both sides (the oracle gets the rounded dY and weights), accumulation is fp32, the
result is stored as bf16: one bf16 ulp of the tensor maximum, as for the forward block (tests/test_gpu_block.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn as nn

from grouped_ssd_pytorch_b200.layers.modules.source_block import PM, _Conv, conv_igemm, dgrad_weight
from oracle.source_block import bf16_round, conv2d_backward


def check(n, c_in, c_out, h, w, groups, k, seed):
    r = np.random.RandomState(seed)
    wt = (r.randn(c_out, c_in // groups, k, k) * np.sqrt(2.0 / (c_in // groups * k * k))).astype(np.float32)
    dy = bf16_round(r.randn(n, c_out, h, w).astype(np.float32))
    x = np.zeros((n, c_in, h, w), np.float32)                                    # only its shape matters for the data gradient
    dx_ref, _, _ = conv2d_backward(x, bf16_round(wt), dy, groups, k // 2)
    # the same gradient as a forward convolution of dY: c_out channels in, c_in channels out, transformed filter, no bias
    conv = nn.Conv2d(c_out, c_in, k, padding=k // 2, groups=groups, bias=False)
    with torch.no_grad():
        conv.weight.copy_(dgrad_weight(torch.from_numpy(wt), groups))
    cv = _Conv(conv.cuda(), groups, dev=torch.device("cuda:0"))
    out = conv_igemm(PM.from_nchw(torch.from_numpy(dy).cuda()), cv, relu=False, shift=cv.shift)
    got = out.to_nchw().cpu().numpy()
    scale = np.abs(dx_ref).max()
    err = np.abs(got - dx_ref).max()
    ok = err <= scale * 2.0 ** -7
    print("n=%d %dx%d  %d -> %d channels, groups %d, %dx%d filter: max |dx - oracle| = %.3e (scale %.3e)  %s"
          % (n, h, w, c_out, c_in, groups, k, k, err, scale, "ok" if ok else "<- WRONG"))
    return ok


if __name__ == "__main__":
    assert torch.cuda.is_available()
    good = True
    good &= check(2, 512, 512, 7, 5, 4, 3, 1)        # vgg.30 (source 1): grouped 3x3
    good &= check(1, 1024, 1024, 5, 5, 4, 1, 2)      # vgg.47 (source 2): grouped 1x1
    good &= check(2, 512, 512, 6, 6, 1, 1, 3)        # fuse_11: dense 1x1
    good &= check(2, 256, 512, 6, 6, 4, 3, 4)        # c_in != c_out: the channel roles really are swapped
    sys.exit(0 if good else 1)
