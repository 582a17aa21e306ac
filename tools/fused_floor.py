"""development: launch + drain floor of the one-launch loss kernel's grid shape (needs the --phase-timing debug build)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grouped_ssd_pytorch_b200 import build
os.environ["GSSD_LIB"] = build.LIB.replace(".so", "_dbg.so")
import torch
from grouped_ssd_pytorch_b200 import _lib, config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import PriorBox
from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list
lib = _lib.require_cuda(); dev = torch.device("cuda:0")
pri = PriorBox(config.v2).forward(device="cuda"); P = pri.shape[0]; B = 32
tg = syn.targets(syn.rng(1), B, 1, 5)
gt, off, sg, gm = pack_target_list([torch.from_numpy(t) for t in tg], dev)
loc = torch.randn(B, P, 4, device=dev); conf = torch.randn(B, P, 2, device=dev)
gl, gc = torch.empty_like(loc), torch.empty_like(conf)
losses = torch.empty(2, device=dev); npos = torch.empty(B, dtype=torch.int32, device=dev)
wsb = lib.gssd_workspace_bytes(_lib.WS_LOSS, B, P, 2, sg, 0); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
state = torch.zeros(int(lib.gssd_fused_state_bytes()), dtype=torch.uint8, device=dev)
st = _lib.stream()
for ratio, label in ((3, "full kernel"), (-12345, "every CTA returns at once")):
    def k():
        _lib.check(lib.gssd_mbox_loss_fused(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, P, 2, gt.data_ptr(), off.data_ptr(), sg, gm, 0.5, ratio, 0.1, 0.2,
                                            state.data_ptr(), None, losses.data_ptr(), gl.data_ptr(), gc.data_ptr(), None, None, npos.data_ptr(), ws.data_ptr(), wsb, st))
    for _ in range(5): k()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300): k()
    e1.record(); torch.cuda.synchronize()
    print("%-28s %.2f us per launch (300 back to back)" % (label, e0.elapsed_time(e1) / 300 * 1e3))
