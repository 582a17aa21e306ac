"""Quick per-kernel timing sweep (CUDA events, rotating inputs > L2) — development aid, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from grouped_ssd_pytorch_b200 import build as _b
if os.environ.get("GSSD_ALT_LIB"):          # development: time an alternative build of the library
    _b.LIB = os.path.abspath(os.environ["GSSD_ALT_LIB"]); _b.stale = lambda: False
from grouped_ssd_pytorch_b200 import _lib, config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import PriorBox
from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list

lib = _lib.require_cuda()
dev = torch.device("cuda:0")
HBM = 6521.4


def timeit(fn, n_sets, iters=int(os.environ.get("QP_ITERS", "20")), warm=3):
    for i in range(warm):
        fn(i % n_sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def run(pname, B, gmax):
    pri = PriorBox(config.ALL[pname]).forward(device="cuda")
    P = pri.shape[0]
    per_set = B * P * 64
    n_sets = max(2, min(12, int(300e6 // per_set) + 1))
    r = syn.rng(1)
    tg = syn.targets(r, B, 1, gmax)
    gt, gt_off, sum_g, g_max = pack_target_list([torch.from_numpy(t) for t in tg], dev)
    locs = [torch.randn(B, P, 4, device=dev) * 0.5 for _ in range(n_sets)]
    confs = [torch.randn(B, P, 2, device=dev) for _ in range(n_sets)]
    scores = [torch.softmax(c + torch.tensor([0.0, -4.0], device=dev), -1) for c in confs]
    tags = torch.empty(B, P, dtype=torch.int16, device=dev)
    stats = torch.empty(16 + 4 * B, dtype=torch.uint8, device=dev)
    losses = torch.empty(2, device=dev)
    gl = [torch.empty_like(l) for l in locs]
    gc = [torch.empty_like(c) for c in confs]
    wsb = lib.gssd_workspace_bytes(_lib.WS_LOSS, B, P, 2, sum_g, 0)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty(B, 2, 200, 5, device=dev)
    import ctypes
    bias = (ctypes.c_float * 2)(0.0, -4.0)
    st = _lib.stream()

    def k_match(i):
        _lib.check(lib.gssd_mbox_match(pri.data_ptr(), P, confs[i].data_ptr(), 2, gt.data_ptr(), gt_off.data_ptr(), B, sum_g, g_max, 0.5, tags.data_ptr(), stats.data_ptr(), st))

    def k_loss(i):
        _lib.check(lib.gssd_mbox_loss(locs[i].data_ptr(), confs[i].data_ptr(), pri.data_ptr(), B, P, 2, gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, tags.data_ptr(), stats.data_ptr(), None, 0, 3, 0.1, 0.2, losses.data_ptr(), gl[i].data_ptr(), gc[i].data_ptr(), None, None, ws.data_ptr(), wsb, st))

    state = torch.zeros(int(lib.gssd_fused_state_bytes()), dtype=torch.uint8, device=dev)
    npos = torch.empty(B, dtype=torch.int32, device=dev)
    fused_ok = lib.gssd_mbox_fused_supported(B, P, 2, g_max) == 1

    def k_fused(i):
        _lib.check(lib.gssd_mbox_loss_fused(locs[i].data_ptr(), confs[i].data_ptr(), pri.data_ptr(), B, P, 2, gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, 0.5, 3, 0.1, 0.2, state.data_ptr(), None, losses.data_ptr(), gl[i].data_ptr(), gc[i].data_ptr(), None, None, npos.data_ptr(), ws.data_ptr(), wsb, st))

    def k_det(i):
        _lib.check(lib.gssd_detect_logits(locs[i].data_ptr(), confs[i].data_ptr(), bias, pri.data_ptr(), B, P, 2, 200, 0.2, 0.45, 0.1, 0.2, out.data_ptr(), None, None, st))

    k_match(0)
    tm, tl, td = timeit(k_match, n_sets), timeit(k_loss, n_sets), timeit(k_det, n_sets)
    tf = timeit(k_fused, n_sets) if fused_ok else float("nan")
    bl, bd = B * P * 64, B * (P * 40 + 8000)
    print("%-8s B=%4d G<=%2d  match %7.1f us | loss %7.1f us (%5.0f GB/s, %4.1f%%) | match+loss %5.1f%% | ONE LAUNCH %7.1f us (%4.1f%%) | detect %7.1f us (%5.0f GB/s, %4.1f%%)" % (
        pname, B, gmax, tm, tl, bl / tl / 1e3, bl / tl / 1e3 / HBM * 100, bl / (tm + tl) / 1e3 / HBM * 100, tf, bl / tf / 1e3 / HBM * 100, td, bd / td / 1e3, bd / td / 1e3 / HBM * 100), flush=True)


if __name__ == "__main__":
    if os.environ.get("QP_ONLY"):             # one configuration, several repeats (A/B runs of library variants)
        for _ in range(3):
            run(os.environ.get("QP_PRIORS", "v2"), int(os.environ["QP_ONLY"]), int(os.environ.get("QP_GMAX", "5")))
        sys.exit(0)
    for pname, B, g in [("v2", 32, 5), ("v2", 64, 5), ("v2", 256, 5), ("v2", 1024, 5), ("v2_512", 64, 32), ("v2_512", 512, 32)]:
        run(pname, B, g)
