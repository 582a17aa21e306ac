"""The attention core of GSSD++'s Self_Attn at the model's sizes: the library's kernels against the reference's torch expression
(permute + bmm + softmax + bmm, fp32) — development aid, CUDA events.   python tools/attn_perf.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grouped_ssd_pytorch_b200.layers.self_attn import attention_core

DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False


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


def ref(theta, phi, g):
    attn = torch.softmax(torch.bmm(theta.permute(0, 2, 1), phi), -1)
    return torch.bmm(g, attn.permute(0, 2, 1)), attn


B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for C, H in ((512, 38), (1024, 19), (512, 10), (256, 5)):
    D, Cv, N = C // 8, C // 2, H * H
    th, ph, g = (torch.randn(B, D, N, device=DEV, requires_grad=True) * 0.5), torch.randn(B, D, N, device=DEV, requires_grad=True), torch.randn(B, Cv, N, device=DEV, requires_grad=True)
    th = th.detach().requires_grad_(True)
    d_o = torch.randn(B, Cv, N, device=DEV)

    def step(fn):
        for t in (th, ph, g):
            t.grad = None
        fn(th, ph, g)[0].backward(d_o)

    with torch.no_grad():
        f_o, f_r = timeit(lambda: attention_core(th, ph, g)), timeit(lambda: ref(th, ph, g))
    s_o, s_r = timeit(lambda: step(attention_core)), timeit(lambda: step(ref))
    gf = 2.0 * B * N * N * (D + Cv) / 1e9
    print("C %4d  %2dx%2d  batch %d: forward ours %7.1f us (%.1f TFLOP/s fp32) torch %7.1f us | fwd+bwd ours %7.1f us torch %7.1f us"
          % (C, H, H, B, f_o, gf / f_o * 1e3, f_r, s_o, s_r))
