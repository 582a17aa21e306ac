"""Where does a HostPipeline step spend its time?  (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grouped_ssd_pytorch_b200 import config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import PriorBox
from grouped_ssd_pytorch_b200.pipeline import HostPipeline

B = 32
pri = PriorBox(config.v2).forward(device="cuda"); P = pri.shape[0]
r = syn.rng(0)
tg = [torch.from_numpy(t) for t in syn.targets(r, B)]
loc = torch.from_numpy(syn.loc(r, B, P)); conf = torch.from_numpy(syn.conf_logits(r, B, P, 2))
for depth in (1, 2, 3, 4):
    for detect in (True, False):
        pipe = HostPipeline(B, pri, depth=depth, conf_thresh=0.2, detect_logits=True, class_bias=(0.0, -4.0), max_gt_rows=B * 5)
        hbs = [pipe.host_buffers() for _ in range(2 * depth)]
        for hb in hbs:
            hb.loc.copy_(loc); hb.conf.copy_(conf); hb.t = None
        def step(i):
            hb = hbs[i % len(hbs)]
            if hb.t is not None: pipe.wait(hb.t)
            hb.t = pipe.submit(hb, tg, detect=detect)
        for i in range(30): step(i)
        for hb in hbs: pipe.wait(hb.t)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 300
        host = 0.0
        for i in range(n):
            t1 = time.perf_counter(); step(i); host += time.perf_counter() - t1
        for hb in hbs: pipe.wait(hb.t)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("depth %d detect %d: %.3f ms/step (host time in submit+wait %.3f ms/step)" % (depth, detect, dt / n * 1e3, host / n * 1e3), flush=True)
        pipe.close()
# raw copies
hb = HostBuffers = None
a = torch.empty(6708228, dtype=torch.uint8).pin_memory(); d = torch.empty_like(a, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for nbytes in (6708228, 4470784, 2235392):
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): d[:nbytes].copy_(a[:nbytes], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D %d bytes: %.3f ms, %.1f GB/s" % (nbytes, e0.elapsed_time(e1) / 50, nbytes / (e0.elapsed_time(e1) / 50 * 1e-3) / 1e9))
# host cost of packing alone
pipe = HostPipeline(B, pri, depth=2, max_gt_rows=B * 5); hb = pipe.host_buffers()
t0 = time.perf_counter()
for _ in range(1000): pipe._pack(hb, tg)
print("pack targets: %.1f us" % ((time.perf_counter() - t0) / 1000 * 1e6))
