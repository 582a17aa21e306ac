"""One forward + backward of GSSD++'s deformable convolution on the library's kernels (target of ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grouped_ssd_pytorch_b200.layers import dcn_v2_custom as ours
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
C, O, H, W, dg = 1024, 512, 38, 38, 4
DEV = "cuda:0"
torch.manual_seed(0)
x = torch.randn(N, C, H, W, device=DEV, requires_grad=True)
w = (torch.randn(O, C, 3, 3, device=DEV) / (9 * C) ** 0.5).requires_grad_(True)
b = torch.zeros(O, device=DEV, requires_grad=True)
off = (float(os.environ.get("DCN_OFF", "1.5")) * torch.randn(N, 2 * dg * 9, H, W, device=DEV)).requires_grad_(True)
msk = torch.sigmoid(torch.randn(N, dg * 9, H, W, device=DEV)).requires_grad_(True)
gout = torch.randn(N, O, H, W, device=DEV)
for _ in range(2):
    y = ours.dcn_v2_conv(x, off, msk, w, b, 1, 1, 1, dg)
    y.backward(gout)
torch.cuda.synchronize()
