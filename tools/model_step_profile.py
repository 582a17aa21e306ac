"""torch.profiler kernel table of the GSSD training step of bench.py's `model_step` (batch 32): python tools/model_step_profile.py [torch|gssd|backbone]"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gssd_standin as G
from grouped_ssd_pytorch_b200 import config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import MultiBoxLoss, PriorBox
from grouped_ssd_pytorch_b200.layers.modules.source_block import gssd_forward
mode = sys.argv[1] if len(sys.argv) > 1 else "gssd"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
net = G.StandInSSD('train', 2, True, PriorBox(config.v2).forward())
net.load_state_dict(G.seeded_state(net.state_dict(), 71)); net.to(dev).train()
x = G.seeded_input(72, B).to(dev)
targets = [torch.from_numpy(t).to(dev) for t in syn.targets(syn.rng(5), B, 1, 5)]
crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True); crit.process_group = False
fast = types.MethodType(gssd_forward, net)
if mode == "backbone":                                       # conv3_2 .. conv5_3 on the tcgen05 kernels too
    fwd = lambda xx: gssd_forward(net, xx, backbone=True)
else:
    fwd = fast if mode == "gssd" else (lambda xx: G.forward_torch(net, xx) + (net.priors,))
def step():
    net.zero_grad(set_to_none=True)
    ll, lc = crit(fwd(x), targets)
    (ll + lc).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
