"""Where a GSSD++ training step (BASELINE configs[2]: reference forward on the drop-in layers + MultiBoxLoss + backward) spends
its time: torch.profiler kernel table + wall clock.   python tools/gssdpp_profile.py [batch] [ref|ours]
`ref`: Self_Attn / DCN stay the reference's torch modules (DCN operator = torchvision); `ours` (default): the drop-in modules."""
import os, sys, time, types, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next(p for p in ("/root/reference/ssd_liverdet", os.path.join(ROOT, "baseline", "_ref", "ssd_liverdet")) if os.path.isdir(p))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, REF)
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "ours"          # ours | ref | dcn (our DCN, the reference's Self_Attn)
mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
import grouped_ssd_pytorch_b200 as gssd
if mode == "ref":
    dcn = types.ModuleType("dcn_v2")
    class _DCNv2:
        @staticmethod
        def apply(inp, off, mask, w, b, stride, pad, dil, dg):
            from torchvision.ops import deform_conv2d
            return deform_conv2d(inp, off, w, b, stride=stride, padding=pad, dilation=dil, mask=mask)
    dcn._DCNv2 = _DCNv2; sys.modules["dcn_v2"] = dcn
    gssd.install_as_layers(reference_modules=("dcn_v2_custom", "self_attn"))
elif mode == "dcn":
    gssd.install_as_layers(reference_modules=("self_attn",))
else:
    gssd.install_as_layers()
from models.ssd_multiphase_custom_group import build_ssd
from layers.modules import MultiBoxLoss
from grouped_ssd_pytorch_b200 import synthetic as syn
net = build_ssd('train', 300, 2, True, 4, 4, 1, True, True, True, 1, 4, True, False, 1)
print("DCN module:", type(net.dcn_list[0]).__module__, "| Self_Attn module:", type(net.self_attn_list[0]).__module__)
torch.manual_seed(3)
net.cuda().train()
x = torch.rand(B, 12, 300, 300, device="cuda")
targets = [torch.from_numpy(t).cuda() for t in syn.targets(syn.rng(9), B, 1, 5)]
criterion = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)
def step():
    net.zero_grad()
    out = net(x)
    ll, lc = criterion(out, targets)
    (ll + lc).backward()
for _ in range(3):
    step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("GSSD++ step, batch %d, %s: %.2f ms (%.0f images/s)" % (B, mode, dt * 1e3, B / dt))
if os.environ.get("NO_PROFILE"):
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
