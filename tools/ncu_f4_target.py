"""One forward + backward of each round-2 addition at its model size (target of ncu captures): GSSD++'s deformable convolution and
Self_Attn core (4 images), the fused BatchNorm + ReLU and the pool backward of the backbone (batch 32, conv1_x size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from grouped_ssd_pytorch_b200.layers import dcn_v2_custom as D, self_attn as S
from grouped_ssd_pytorch_b200.layers.modules.bn_relu import bn_relu, max_pool
DEV = "cuda:0"
torch.manual_seed(0)
N, C, O, H, dg = 4, 1024, 512, 38, 4
x = torch.randn(N, C, H, H, device=DEV, requires_grad=True)
w = (torch.randn(O, C, 3, 3, device=DEV) / (9 * C) ** 0.5).requires_grad_(True)
b = torch.zeros(O, device=DEV, requires_grad=True)
off = (1.5 * torch.randn(N, 2 * dg * 9, H, H, device=DEV)).requires_grad_(True)
msk = torch.sigmoid(torch.randn(N, dg * 9, H, H, device=DEV)).requires_grad_(True)
for _ in range(2):
    D.dcn_v2_conv(x, off, msk, w, b, 1, 1, 1, dg).backward(torch.randn(N, O, H, H, device=DEV))
th = (0.5 * torch.randn(N, 64, H * H, device=DEV)).requires_grad_(True)
ph = torch.randn(N, 64, H * H, device=DEV, requires_grad=True)
g = torch.randn(N, 256, H * H, device=DEV, requires_grad=True)
for _ in range(2):
    S.attention_core(th, ph, g)[0].backward(torch.randn(N, 256, H * H, device=DEV))
xb = torch.randn(32, 64, 300, 300, device=DEV, requires_grad=True)
bn = nn.BatchNorm2d(64).to(DEV).train()
pool = nn.MaxPool2d(2, 2)
gout = torch.randn(32, 64, 150, 150, device=DEV)
for _ in range(2):
    xb.grad = None
    max_pool(bn_relu(xb, bn), pool).backward(gout)
torch.cuda.synchronize()
