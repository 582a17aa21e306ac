"""Which torch ops surround the two loss kernels in one training step (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from grouped_ssd_pytorch_b200 import config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import PriorBox, MultiBoxLoss
from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list

dev = torch.device("cuda:0")
pri = PriorBox(config.v2).forward(device="cuda")
B, P = 32, pri.shape[0]
r = syn.rng(1)
tg = pack_target_list([torch.from_numpy(t) for t in syn.targets(r, B, 1, 5)], dev)
loc = (torch.randn(B, P, 4, device=dev) * 0.5).requires_grad_()
conf = torch.randn(B, P, 2, device=dev).requires_grad_()
crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)


def step():
    loc.grad = None; conf.grad = None
    ll, lc = crit((loc, conf, pri), tg)
    (ll + lc).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::") and "fill" in e.name or e.name in ("aten::add", "aten::ones_like", "aten::zeros", "aten::select"):
        print(e.name, [str(s) for s in (e.stack or [])[:6]], [str(i) for i in (e.input_shapes or [])])
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
