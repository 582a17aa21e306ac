"""Launch the source-1 grouped conv / fuse / heads kernels a few times (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from grouped_ssd_pytorch_b200.layers.modules.source_block import PM, _Conv, conv_igemm
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
x = PM.from_nchw(torch.relu(torch.randn(B, 512, 38, 38, device=dev)))
g = _Conv(nn.Conv2d(512, 512, 3, padding=1, groups=4).to(dev), 4, dev=dev)
f = _Conv(nn.Conv2d(512, 512, 1).to(dev), 1, dev=dev)
h = _Conv(nn.Conv2d(512, 16, 3, padding=1).to(dev), 1, extra=nn.Conv2d(512, 8, 3, padding=1).to(dev), dev=dev)
P = 38 * 38 * 4
loc, conf = torch.empty(B, P, 4, device=dev), torch.empty(B, P, 2, device=dev)
for _ in range(3):
    y = conv_igemm(x, g, relu=True, shift=g.bias)
    s = conv_igemm(y, f, relu=True, shift=f.bias)
    conv_igemm(s, h, relu=False, shift=h.bias, head=(loc, conf, 4, 2, 0, P))
torch.cuda.synchronize()
print("done")
