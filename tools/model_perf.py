"""GSSD forward at batch B: the model's own torch forward (fp32, cuDNN) vs gssd_forward with the tcgen05 source blocks
(and, with backbone=True, conv3_2 .. conv5_3 in PM/bf16 as well).  Development aid: builds the stand-in model of tests/."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gssd_standin as G
from grouped_ssd_pytorch_b200 import config
from grouped_ssd_pytorch_b200.layers import PriorBox
from grouped_ssd_pytorch_b200.layers.modules.source_block import gssd_forward

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.backends.cudnn.benchmark = True
net = G.StandInSSD('train', 2, True, PriorBox(config.v2).forward())
net.load_state_dict(G.seeded_state(net.state_dict(), 71))
net.eval().cuda()
x = torch.rand(B, 12, 300, 300, device="cuda")


def timeit(fn, iters=10):
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    ms = timeit(lambda: G.forward_torch(net, x))
    print("torch forward (cuDNN fp32%s)          : %7.3f ms  %8.0f img/s" % (", TF32 allowed" if tf32 else "", ms, B / ms * 1e3), flush=True)
torch.backends.cudnn.allow_tf32 = False
for bb in (False, True):
    ms = timeit(lambda: gssd_forward(net, x, backbone=bb))
    print("gssd_forward backbone=%-5s             : %7.3f ms  %8.0f img/s" % (bb, ms, B / ms * 1e3), flush=True)
