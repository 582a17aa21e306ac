"""development: stage-by-stage comparison of the source-block backward with torch autograd on the same modules (GPU, fp32)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.nn.functional as F
import cases
from test_gpu_block import modules_from
from grouped_ssd_pytorch_b200.layers import SourceBlock

def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))

for tag in sys.argv[1:] or ["s4", "s1_nobn", "s1_train"]:
    x, prm, training = cases.block_case(tag)
    seed, N, C, H, W, gc, bn, l2, Cf, A, ncls, _ = cases.BLOCK_CASES[tag]
    gconv, gbn, l2m, fuse, bnf, locm, confm = modules_from(tag, prm)
    blk = SourceBlock(gconv, gbn, l2m, fuse, bnf, locm, confm, num_classes=ncls)
    blk._debug_backward = {}
    xt = torch.from_numpy(x).cuda().requires_grad_()
    loc, conf, xo = blk.forward_autograd(xt)
    d_loc, d_conf = cases.block_upstream(tag, loc[0].numel(), conf[0].numel())
    T = lambda a: torch.from_numpy(np.asarray(a, np.float32)).cuda()
    ((loc.reshape(N, -1) * T(d_loc)).sum() + (conf.reshape(N, -1) * T(d_conf)).sum()).backward()
    D = blk._debug_backward
    # torch reference with retained intermediates
    xr = torch.from_numpy(x).cuda().requires_grad_()
    h = xr
    if gconv is not None:
        h = gconv(h)
        if gbn is not None: h = gbn(h)
        h = F.relu(h)
    h.retain_grad()
    s = l2m(h) if l2m is not None else h
    s.retain_grad()
    zraw = fuse(s); zraw.retain_grad()
    z = F.relu(bnf(zraw)) if bnf is not None else F.relu(zraw)
    z.retain_grad()
    lo = locm(z).permute(0, 2, 3, 1).reshape(N, -1); co = confm(z).permute(0, 2, 3, 1).reshape(N, -1)
    ((lo * T(d_loc)).sum() + (co * T(d_conf)).sum()).backward()
    print("==", tag, "training", training)
    print("  forward: z2 vs torch %.2e   y1 vs torch %.2e" % (rel(D["z2"], z.detach()), rel(D["y1"], h.detach())))
    print("  dz2 (grad wrt relu(bn_fuse(.))) %.2e" % rel(D["dz2"], z.grad))
    flips = ((D["z2"] > 0) != (z.detach() > 0)).sum().item()
    print("  ReLU masks that differ from torch's (bf16 forward, pre-activation near 0): %d of %d" % (flips, z.numel()))
    same = ((D["z2"] > 0) == (z.detach() > 0)).float()
    if l2m is None:
        print("  t2  (grad wrt fuse output)      %.2e   where the masks agree: %.2e" % (rel(D["t2"], zraw.grad), rel(D["t2"] * same, zraw.grad * same)))
        print("  a1  (grad wrt fuse input)       %.2e" % rel(D["a1"], s.grad))
    else:
        rs = 1.0 / (D["ss"].view(N, H + 2, W + 2)[:, 1:-1, 1:-1].sqrt() + 1e-10)
        print("  t2/rs (grad wrt fuse output)    %.2e   where the masks agree: %.2e" % (rel(D["t2"] / rs.unsqueeze(1), zraw.grad), rel(D["t2"] / rs.unsqueeze(1) * same, zraw.grad * same)))
        print("  a1 (= dz*l2w/r, before the L2 correction) vs s.grad*l2w*rs: %.2e" % rel(D["a1"], s.grad * l2m.weight.view(1, -1, 1, 1) * rs.unsqueeze(1)))
    if gconv is not None:
        print("  x.grad %.2e" % rel(xt.grad, xr.grad))
    else:
        print("  x.grad %.2e" % rel(xt.grad, xr.grad))
    print("  fuse.weight.grad: ours vs torch (same modules -> grads accumulated twice; compare halves)")
