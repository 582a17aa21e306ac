"""Host-side cost breakdown of one e2e step (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from grouped_ssd_pytorch_b200 import _lib, config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import Detect, MultiBoxLoss, PriorBox
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pri = PriorBox(config.v2).forward(device="cuda"); P = pri.shape[0]
r = syn.rng(0)
tg = [torch.from_numpy(t) for t in syn.targets(r, B)]
loc_h = torch.from_numpy(syn.loc(r, B, P)).pin_memory(); conf_h = torch.from_numpy(syn.conf_logits(r, B, P)).pin_memory()
sc_h = torch.from_numpy(syn.detect_scores(r, B, P)).pin_memory()
crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
out_h = torch.empty(B, 2, 200, 5).pin_memory(); loss_h = torch.empty(2).pin_memory()
T = {}
def tick(name, t0, sync=True):
    if sync: torch.cuda.synchronize()
    T[name] = T.get(name, 0) + time.perf_counter() - t0
N = 200
for it in range(N + 20):
    if it == 20: T.clear()
    t = time.perf_counter(); loc = loc_h.to(dev, non_blocking=True).requires_grad_(); conf = conf_h.to(dev, non_blocking=True).requires_grad_(); sc = sc_h.to(dev, non_blocking=True); tick("h2d", t)
    t = time.perf_counter(); ll, lc = crit((loc, conf, pri), tg); tick("loss_fwd", t)
    t = time.perf_counter(); (ll + lc).backward(); tick("backward", t)
    t = time.perf_counter(); out = Detect.apply(2, 0, 200, 0.2, 0.45, loc.detach(), sc, pri); tick("detect", t)
    t = time.perf_counter(); loss_h.copy_(torch.stack([ll.detach(), lc.detach()]), non_blocking=True); out_h.copy_(out, non_blocking=True); tick("d2h", t)
print("B=%d per-step host+device ms (each sub-step synchronised):" % B, {k: round(v / N * 1e3, 4) for k, v in T.items()}, "sum", round(sum(T.values()) / N * 1e3, 4))
# without intermediate syncs
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(N):
    loc = loc_h.to(dev, non_blocking=True).requires_grad_(); conf = conf_h.to(dev, non_blocking=True).requires_grad_(); sc = sc_h.to(dev, non_blocking=True)
    ll, lc = crit((loc, conf, pri), tg); (ll + lc).backward()
    out = Detect.apply(2, 0, 200, 0.2, 0.45, loc.detach(), sc, pri)
    loss_h.copy_(torch.stack([ll.detach(), lc.detach()]), non_blocking=True); out_h.copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
print("one sync per step: %.4f ms/step" % ((time.perf_counter() - t0) / N * 1e3))
# host-only cost of the API calls (GPU work queued, no sync)
import cProfile, pstats
loc = loc_h.to(dev).requires_grad_(); conf = conf_h.to(dev).requires_grad_(); sc = sc_h.to(dev)
def body():
    for _ in range(100):
        ll, lc = crit((loc, conf, pri), tg); (ll + lc).backward()
        Detect.apply(2, 0, 200, 0.2, 0.45, loc.detach(), sc, pri)
    torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); body(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
