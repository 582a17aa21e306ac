"""Generate tests/golden/self_attn.npz: outputs and gradients of the reference's OWN `Self_Attn` module
(layers/self_attn.py:29-89, spectral norm from layers/spectral_norm.py) on the seeded cases of tests/cases.py (SA_CASES), in
float64 on the CPU, evaluation mode (no power iteration, so the result is a function of the stored parameters alone).  The pin of
oracle/self_attn.py and the golden of tests/test_gpu_self_attn.py.

Runs only in the build container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_self_attn.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/ssd_liverdet")
warnings.filterwarnings("ignore")

import torch  # noqa: E402
from layers.self_attn import Self_Attn  # noqa: E402  (reference)

import cases  # noqa: E402

torch.set_num_threads(4)


def main():
    out = {}
    for tag, (seed, B, C, H, factor) in cases.SA_CASES.items():
        x, prm, (u1, u2) = cases.sa_case(tag)
        m = Self_Attn(C, factor).double().eval()
        missing = m.load_state_dict({k: torch.from_numpy(np.asarray(v)).double() for k, v in prm.items()}, strict=True)
        xt = torch.from_numpy(x).double().requires_grad_(True)
        kept = {}

        def keep(name):
            def hook(mod, inputs, o):
                o.retain_grad()
                kept[name] = o
            return hook

        def keep_in(mod, inputs):
            inputs[0].retain_grad()
            kept["attn_g"] = inputs[0]

        hooks = [getattr(m, "snconv1x1_" + n).register_forward_hook(keep(n)) for n in ("theta", "phi", "g")]
        hooks.append(m.snconv1x1_attn.register_forward_pre_hook(keep_in))
        y, gated, attn = m(xt, True)
        for hk in hooks:
            hk.remove()
        ((y * torch.from_numpy(u1).double()).sum() + (gated * torch.from_numpy(u2).double()).sum()).backward()
        out[tag + "_out"], out[tag + "_gated"], out[tag + "_attn"] = y.detach().numpy(), gated.detach().numpy(), attn.detach().numpy()
        # what flows through the attention core (self_attn.py:69-81) inside the reference's own backward: the gradient of attn_g
        # (input of snconv1x1_attn) and of the theta / phi / g convolutions' outputs (phi / g: before the pooling)
        out[tag + "_d_attn_g"] = kept["attn_g"].grad.numpy()
        for n in ("theta", "phi", "g"):
            out[tag + "_d_" + n + "_conv"] = kept[n].grad.numpy()
        out[tag + "_d_x"] = xt.grad.numpy()
        out[tag + "_d_sigma"] = m.sigma.grad.numpy()
        for name in ("snconv1x1_theta", "snconv1x1_phi", "snconv1x1_g", "snconv1x1_attn"):
            out["%s_d_%s_w_s" % (tag, name)] = cases.strided_sample(getattr(m, name).weight_orig.grad.numpy())
            out["%s_d_%s_b" % (tag, name)] = getattr(m, name).bias.grad.numpy()
        print(tag, "out", tuple(y.shape), "attn", tuple(attn.shape), "attn row max %.3f" % float(attn.max()), missing)
    out = {k: (v if k.endswith("_s") else v.astype(np.float32)) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "self_attn.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(HERE, "self_attn.npz")), "bytes")


if __name__ == "__main__":
    main()
