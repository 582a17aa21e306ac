"""GSSD++'s deformable convolution (1024 -> 512 channels, 38 x 38, 4 deformable groups) on the library's kernels against
torchvision's fp32 CUDA operator — development aid, CUDA events.   python tools/dcn_perf.py [batch ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchvision.ops import deform_conv2d
from grouped_ssd_pytorch_b200 import _lib
from grouped_ssd_pytorch_b200.layers import dcn_v2_custom as ours
from grouped_ssd_pytorch_b200.layers.modules.source_block import PM

DEV = "cuda:0"
lib = _lib.require_cuda()


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def run(N, C=1024, O=512, H=38, W=38, dg=4):
    x = torch.randn(N, C, H, W, device=DEV)
    w = (torch.randn(O, C, 3, 3, device=DEV) / (9 * C) ** 0.5).requires_grad_(True)
    b = torch.zeros(O, device=DEV, requires_grad=True)
    off = (float(os.environ.get("DCN_OFF", "1.5")) * torch.randn(N, 2 * dg * 9, H, W, device=DEV)).requires_grad_(True)   # std of the offsets, pixels
    msk = torch.sigmoid(torch.randn(N, dg * 9, H, W, device=DEV)).requires_grad_(True)
    gout = torch.randn(N, O, H, W, device=DEV)
    xg = x.clone().requires_grad_(True)

    def step(fn):
        for t in (xg, w, b, off, msk):
            t.grad = None
        y = fn()
        y.backward(gout)

    f_ours = lambda: ours.dcn_v2_conv(xg, off, msk, w, b, 1, 1, 1, dg)
    f_tv = lambda: deform_conv2d(xg, off, w, b, stride=1, padding=1, dilation=1, mask=msk)
    with torch.no_grad():
        t_of, t_tf = timeit(f_ours), timeit(f_tv, iters=3, warm=1)
    t_o, t_t = timeit(lambda: step(f_ours)), timeit(lambda: step(f_tv), iters=3, warm=1)
    # the two gather / scatter kernels alone
    xp = PM.from_nchw(x)
    col = PM.empty(N, 9 * C, H, W, x.device)
    st = _lib.stream()
    offd, mskd = off.detach(), msk.detach()
    t_col = timeit(lambda: _lib.check(lib.gssd_dcn_columns(xp.data.data_ptr(), offd.data_ptr(), mskd.data_ptr(), N, C, H, W, dg, col.data.data_ptr(), st)))
    dx = torch.empty(xp.rows, C, device=DEV); do, dm = torch.empty_like(offd), torch.empty_like(mskd)
    t_bwd = timeit(lambda: _lib.check(lib.gssd_dcn_columns_bwd(xp.data.data_ptr(), offd.data_ptr(), mskd.data_ptr(), col.data.data_ptr(), N, C, H, W, dg,
                                                               dx.data_ptr(), do.data_ptr(), dm.data_ptr(), st)))
    flops = 2.0 * N * H * W * O * 9 * C
    col_bytes = xp.rows * 9 * C * 2
    print("batch %3d: forward ours %8.1f us (%.0f TFLOP/s algorithmic) torchvision %9.1f us (x%.1f) | fwd+bwd ours %8.1f us torchvision %9.1f us (x%.1f)"
          % (N, t_of, flops / t_of / 1e6, t_tf, t_tf / t_of, t_o, t_t, t_t / t_o))
    print("           gssd_dcn_columns %7.1f us (%.0f GB/s of column writes)   gssd_dcn_columns_bwd %7.1f us" % (t_col, col_bytes / t_col / 1e3, t_bwd))


for n in [int(a) for a in sys.argv[1:]] or [4, 32]:
    run(n)
