"""Launch each hot-path kernel a few times at a given batch (target of ncu captures)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grouped_ssd_pytorch_b200 import _lib, config, synthetic as syn
from grouped_ssd_pytorch_b200.layers import PriorBox
from grouped_ssd_pytorch_b200.layers.box_utils import pack_target_list

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pname = sys.argv[2] if len(sys.argv) > 2 else "v2"
gmax = int(sys.argv[3]) if len(sys.argv) > 3 else 5
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
lib = _lib.require_cuda()
dev = torch.device("cuda:0")
pri = PriorBox(config.ALL[pname]).forward(device="cuda")
P = pri.shape[0]
r = syn.rng(1)
tg = syn.targets(r, B, 1, gmax)
gt, gt_off, sum_g, g_max = pack_target_list([torch.from_numpy(t) for t in tg], dev)
loc = torch.randn(B, P, 4, device=dev) * 0.5
conf = torch.randn(B, P, 2, device=dev)
tags = torch.empty(B, P, dtype=torch.int16, device=dev)
stats = torch.empty(16 + 4 * B, dtype=torch.uint8, device=dev)
losses = torch.empty(2, device=dev)
gl, gc = torch.empty_like(loc), torch.empty_like(conf)
wsb = lib.gssd_workspace_bytes(_lib.WS_LOSS, B, P, 2, sum_g, 0)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
out = torch.empty(B, 2, 200, 5, device=dev)
bias = (ctypes.c_float * 2)(0.0, -4.0)
st = _lib.stream()
state = torch.zeros(int(lib.gssd_fused_state_bytes()), dtype=torch.uint8, device=dev)
npos = torch.empty(B, dtype=torch.int32, device=dev)
fused_ok = lib.gssd_mbox_fused_supported(B, P, 2, g_max) == 1
for _ in range(reps):
    if fused_ok:
        _lib.check(lib.gssd_mbox_loss_fused(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, P, 2, gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, 0.5, 3, 0.1, 0.2, state.data_ptr(), None, losses.data_ptr(), gl.data_ptr(), gc.data_ptr(), None, None, npos.data_ptr(), ws.data_ptr(), wsb, st))
    _lib.check(lib.gssd_mbox_match(pri.data_ptr(), P, conf.data_ptr(), 2, gt.data_ptr(), gt_off.data_ptr(), B, sum_g, g_max, 0.5, tags.data_ptr(), stats.data_ptr(), st))
    _lib.check(lib.gssd_mbox_loss(loc.data_ptr(), conf.data_ptr(), pri.data_ptr(), B, P, 2, gt.data_ptr(), gt_off.data_ptr(), sum_g, g_max, tags.data_ptr(), stats.data_ptr(), None, 0, 3, 0.1, 0.2, losses.data_ptr(), gl.data_ptr(), gc.data_ptr(), None, None, ws.data_ptr(), wsb, st))
    _lib.check(lib.gssd_detect_logits(loc.data_ptr(), conf.data_ptr(), bias, pri.data_ptr(), B, P, 2, 200, 0.2, 0.45, 0.1, 0.2, out.data_ptr(), None, None, st))
torch.cuda.synchronize()
print("done", losses.tolist())
