"""Time the tcgen05 implicit-GEMM convolutions of the source block at configs[1] size (development aid; the
numbers that count are taken by bench.py).

    python tools/gconv_perf.py [batch]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from grouped_ssd_pytorch_b200.layers.modules.source_block import PM, _Conv, conv_igemm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
PEAK = 1651.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass


def timeit(fn, iters=20, warm=3):
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


def case(name, c_in, c_out, k, groups, hw, head=None):
    conv = nn.Conv2d(c_in, c_out, k, padding=(k - 1) // 2, groups=groups).to(dev)
    x = PM.from_nchw(torch.relu(torch.randn(B, c_in, hw, hw, device=dev)))
    if head is None:
        cv = _Conv(conv, groups, dev=dev)
        fn = lambda: conv_igemm(x, cv, relu=True, shift=cv.bias)
    else:
        A, ncls = head
        conf = nn.Conv2d(c_in, A * ncls, k, padding=1).to(dev)
        cv = _Conv(conv, 1, extra=conf, dev=dev)
        P = hw * hw * A
        loc_o, conf_o = torch.empty(B, P, 4, device=dev), torch.empty(B, P, ncls, device=dev)
        fn = lambda: conv_igemm(x, cv, relu=False, shift=cv.bias, head=(loc_o, conf_o, A, ncls, 0, P))
        c_out = cv.c_out
    us = timeit(fn)
    if os.environ.get("GCONV_DBG"):
        import ctypes
        from grouped_ssd_pytorch_b200 import _lib
        lib = _lib.load()
        buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        lib.gssd_debug_conv_timing(ctypes.c_void_p(buf.data_ptr()))
        lib.gssd_debug_conv_flags(int(os.environ.get("GCONV_FLAGS", "0")))
        fn(); torch.cuda.synchronize()
        lib.gssd_debug_conv_flags(0)
        lib.gssd_debug_conv_timing(ctypes.c_void_p(0))
        b = buf.view(148, 16).cpu().double()
        b = b[b[:, 6] > 0]
        names = ["prod wait A-empty", "prod wait B-empty", "prod total", "mma wait T-empty", "mma wait A-full", "mma wait B-full", "mma total", "epi wait T-full", "epi total", "mma issue", "mma commit", "epi tmem-ld", "epi store-buf wait"]
        print("    " + " | ".join("%s %.0f" % (n, b[:, i].mean()) for i, n in enumerate(names)) + " | ctas %d max mma total %.0f" % (b.shape[0], b[:, 6].max()), flush=True)
    flops = 2.0 * B * hw * hw * c_out * (c_in // groups) * k * k
    padded = 2.0 * B * (hw + 2) * (hw + 2) * c_out * (c_in // groups) * k * k
    # cuDNN comparison (library, informational)
    xt = torch.relu(torch.randn(B, c_in, hw, hw, device=dev)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    cb = conv.to(torch.bfloat16).to(memory_format=torch.channels_last)
    us_ref = float("nan")
    if not os.environ.get("GCONV_NOCUDNN"):
        with torch.no_grad():
            us_ref = timeit(lambda: cb(xt))
    print("%-28s %8.1f us  %7.1f TFLOP/s algorithmic (%4.1f%% of %.0f), %7.1f issued | cuDNN bf16 NHWC %8.1f us" % (
        name, us, flops / us / 1e6, flops / us / 1e6 / PEAK * 100, PEAK, padded / us / 1e6, us_ref), flush=True)


if __name__ == "__main__":
    torch.backends.cudnn.benchmark = True
    if os.environ.get("GCONV_ONLY"):
        case("src1 vgg.30 3x3 g4 512", 512, 512, 3, 4, 38)
        case("src1 fuse_11 1x1 512", 512, 512, 1, 1, 38)
        case("src1 heads 3x3 512->24", 512, 16, 3, 1, 38, head=(4, 2))
        sys.exit(0)
    case("src1 vgg.30 3x3 g4 512", 512, 512, 3, 4, 38)
    case("src1 fuse_11 1x1 512", 512, 512, 1, 1, 38)
    case("src1 heads 3x3 512->24", 512, 16, 3, 1, 38, head=(4, 2))
    case("src2 vgg.47 1x1 g4 1024", 1024, 1024, 1, 4, 19)
    case("src2 fuse_21 1x1 1024", 1024, 1024, 1, 1, 19)
    case("src2 heads 3x3 1024->36", 1024, 24, 3, 1, 19, head=(6, 2))
    case("src3 fuse_31 1x1 512", 512, 512, 1, 1, 10)
    case("src3 heads 3x3 512->36", 512, 24, 3, 1, 10, head=(6, 2))
