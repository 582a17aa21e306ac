"""Generate tests/golden/dcn.npz: outputs and gradients of the reference's OWN `DCN` module (layers/dcn_v2_custom.py:58-88) on the
seeded cases of tests/cases.py (DCN_CASES), in float64 on the CPU.  The module's operator, `dcn_v2._DCNv2.apply`, is a third-party
CUDA extension the reference does not vendor; it is routed to `torchvision.ops.deform_conv2d` — the same operator (modulated
deformable convolution, DCNv2 argument layout), the stand-in SURVEY App. A names.  This is the pin of oracle/dcn.py and the golden
of the GPU tests of gssd_dcn_columns / gssd_dcn_columns_bwd (tests/test_gpu_dcn.py).

Runs only in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_dcn.py
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/ssd_liverdet")
warnings.filterwarnings("ignore")

import torch  # noqa: E402
from torchvision.ops import deform_conv2d  # noqa: E402

dcn = types.ModuleType("dcn_v2")


class _DCNv2:
    @staticmethod
    def apply(inp, off, mask, w, b, stride, pad, dil, dg):
        return deform_conv2d(inp, off, w, b, stride=stride, padding=pad, dilation=dil, mask=mask)


dcn._DCNv2 = _DCNv2
sys.modules["dcn_v2"] = dcn

from layers.dcn_v2_custom import DCN  # noqa: E402  (reference)

import cases  # noqa: E402

torch.set_num_threads(4)
T = lambda a: torch.from_numpy(np.asarray(a)).double()


def main():
    out = {}
    for tag, (seed, N, C, O, H, W, dg, s) in cases.DCN_CASES.items():
        c = cases.dcn_case(tag)
        m = DCN(C, O, kernel_size=3, stride=1, padding=1, deformable_groups=dg).double()
        with torch.no_grad():
            m.weight.copy_(T(c["weight"])); m.bias.copy_(T(c["bias"]))
            m.conv_offset_mask.weight.copy_(T(c["com_w"])); m.conv_offset_mask.bias.copy_(T(c["com_b"]))
        x = T(c["x"]).requires_grad_(True)
        kept = {}

        def keep(mod, inputs, o):
            o.retain_grad()
            kept["om"] = o

        h = m.conv_offset_mask.register_forward_hook(keep)
        y, offset = m(x)
        h.remove()
        (y * T(c["gout"])).sum().backward()
        om = kept["om"]
        n_out = int(((offset.detach().abs() > 0).sum()))
        out[tag + "_out"] = y.detach().numpy()
        out[tag + "_offset"] = offset.detach().numpy()
        out[tag + "_om"] = om.detach().numpy()                    # raw output of conv_offset_mask: (o1 | o2 | mask logits)
        out[tag + "_d_om"] = om.grad.numpy()                       # = (d_offset | d_mask * sigmoid'(logit))
        out[tag + "_d_input"] = x.grad.numpy()
        out[tag + "_d_weight_s"] = cases.strided_sample(m.weight.grad.numpy())     # every 5th element + sum + abs sum
        out[tag + "_d_bias"] = m.bias.grad.numpy()
        print(tag, "out", tuple(y.shape), "abs max %.3f" % float(y.abs().max()), "offsets |max| %.2f" % float(offset.abs().max()),
              "nonzero offsets", n_out)
    out = {k: (v if k.endswith("_s") else v.astype(np.float32)) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "dcn.npz"), **out)
    print("wrote", os.path.join(HERE, "dcn.npz"), os.path.getsize(os.path.join(HERE, "dcn.npz")), "bytes")


if __name__ == "__main__":
    main()
