"""development: parameter gradients of the reference GSSD through gssd_forward vs through its own forward, fixed upstream gradient"""
import os, sys, types, warnings
warnings.filterwarnings("ignore")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next(p for p in ("/root/reference/ssd_liverdet", os.path.join(ROOT, "baseline", "_ref", "ssd_liverdet")) if os.path.isdir(p))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, REF)
dcn = types.ModuleType("dcn_v2"); dcn._DCNv2 = type("_DCNv2", (), {"apply": staticmethod(lambda *a: None)}); sys.modules["dcn_v2"] = dcn
mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
import grouped_ssd_pytorch_b200 as gssd
gssd.install_as_layers()
from models.ssd_multiphase_custom_group import build_ssd
import gssd_standin as G
from grouped_ssd_pytorch_b200.layers.modules.source_block import gssd_forward
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = build_ssd('train', 300, 2, True, 4, 4, 1, True, False, False, 0, 1, False, False, 1).cuda()
xb = G.seeded_input(73, B).cuda()
fast = types.MethodType(gssd_forward, net)
offs = [0, 5776, 7942, 8542, 8692, 8728, 8732]
torch.manual_seed(11)
u, v = torch.randn(B, 8732, 4, device="cuda"), torch.randn(B, 8732, 2, device="cuda")
for mode in ("eval", "train"):
    for n_src in (1, 2, 3, 6):
        uu, vv = u.clone(), v.clone()
        uu[:, offs[n_src]:] = 0; vv[:, offs[n_src]:] = 0
        def grads_of(fwd):
            net.zero_grad()
            net.load_state_dict(G.seeded_state(net.state_dict(), 71))
            net.train(mode == "train")
            o = fwd(xb)
            ((o[0] * uu).sum() + (o[1] * vv).sum()).backward()
            return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        a, b = grads_of(net.__call__), grads_of(fast)
        skip = {n + ".bias" for n, m in net.named_modules() if isinstance(m, torch.nn.Conv2d) and not n.startswith(("loc.", "conf."))} if mode == "train" else set()
        errs = {n: float((b[n] - a[n]).norm() / (a[n].norm() + 1e-30)) for n in a if n in b and n not in skip and float(a[n].norm()) > 0}
        groups = {}
        for n, e in errs.items():
            key = n.split(".")[0] + ("" if not n.startswith("vgg") else (".early" if int(n.split(".")[1]) < 30 else ".late"))
            groups.setdefault(key, []).append(e)
        print(mode, "upstream on sources 1..%d:" % n_src, {k: "%.1e" % float(np.median(v_)) for k, v_ in sorted(groups.items())}, "max %.1e" % max(errs.values()), flush=True)
