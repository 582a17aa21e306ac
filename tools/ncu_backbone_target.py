"""One forward + backward of a grouped backbone triple (Conv2d(groups=4) + BatchNorm2d + ReLU) on the tcgen05 path (PMConvLayer) at the
model's batch-32 sizes — conv3_2 (256 -> 256 at 75 x 75, 64 channels per group) and conv4_2 (512 -> 512 at 38 x 38, 128 per group) —
as the target of an ncu capture; only the second iteration of each is inside cudaProfilerStart / Stop:

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:"conv_igemm|wgrad_kernel|bn_bwd_pm|bn_act_pm|nchw_to_pm|pm_to_nchw" -o gpurun_out/bb python tools/ncu_backbone_target.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from grouped_ssd_pytorch_b200.layers.modules.bn_relu import run_layers
DEV = "cuda:0"
torch.manual_seed(0)
for (c_in, c_out, hw) in ((256, 256, 75), (512, 512, 38)):
    mods = nn.ModuleList([nn.Conv2d(c_in, c_out, 3, padding=1, groups=4), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True)]).to(DEV).train()
    x = torch.randn(32, c_in, hw, hw, device=DEV).relu().requires_grad_()
    gout = torch.randn(32, c_out, hw, hw, device=DEV)
    cache = {}
    for it in range(2):
        x.grad = None
        mods.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.profiler.start()
        run_layers(mods, x, tc=cache).backward(gout)
        torch.cuda.synchronize()
        if it == 1:
            torch.cuda.profiler.stop()
