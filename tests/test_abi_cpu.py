"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/gssd.h declares; the host layer mirrors the reference's module tree and error behaviour and
refuses to run without a CUDA device (no CPU fallback).  No compute call is made here."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import pytest
import torch

import cases
import grouped_ssd_pytorch_b200 as pkg
from grouped_ssd_pytorch_b200 import _lib, build, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gssd.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"GSSD_API[^;(]*?\b(gssd_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ("gssd_priorbox", "gssd_jaccard", "gssd_match", "gssd_encode", "gssd_decode", "gssd_nms",
                 "gssd_mbox_match", "gssd_mbox_loss", "gssd_detect", "gssd_l2norm_fwd", "gssd_log_sum_exp"):
        assert must in syms
    assert len(syms) >= 20


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), "libgssd_b200.so does not export " + s
    assert lib.gssd_abi_version() == 2
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    exported = set(re.findall(r" T (\w+)", out))
    extra = {e for e in exported if not e.startswith("gssd_")}
    assert not extra, "unexpected exports: %s" % sorted(extra)[:5]


def test_binding_table_matches_header():
    assert sorted(_lib._SIGS) == declared_symbols()
    lib = _lib.load()
    assert lib.gssd_error_string(-4).decode().startswith("value error")
    assert lib.gssd_stats_bytes(32) == 16 + 4 * 32
    assert lib.gssd_workspace_bytes(_lib.WS_LOSS, 32, 8732, 2, 100, 0) >= 32 * 8 * 16


def test_sm100a_code_is_in_the_library():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_priorbox_count_host_side():
    lib = _lib.load()
    expect = {"v2": 8732, "v2_512": 24564, "v2_custom": 11620, "v2_custom_512": 32756,
              "v2_custom_squareonly": 8732, "v1": 7308}
    for name, p in expect.items():
        assert lib.gssd_priorbox_count(_lib.prior_cfg(config.ALL[name])) == p
    bad = dict(config.v2); bad["variance"] = [0.1, -1]
    assert lib.gssd_priorbox_count(_lib.prior_cfg(bad)) == _lib.ERR_VALUE


def test_argument_errors_are_reported_before_any_launch():
    lib = _lib.load()
    assert lib.gssd_detect(None, None, None, 1, 1, 2, 200, 0.1, 0.45, 0.1, 0.2, None, None, None, None) == _lib.ERR_ARG
    assert lib.gssd_detect(1, 1, 1, 1, 10, 2, 200, 0.1, 0.0, 0.1, 0.2, 1, None, None, None) == _lib.ERR_VALUE
    assert lib.gssd_detect(1, 1, 1, 1, 10 ** 6, 2, 200, 0.1, 0.45, 0.1, 0.2, 1, None, None, None) == _lib.ERR_LIMIT
    assert lib.gssd_match(1, 100, 1, 1, 2, 0, 0, 0.5, 0.1, 0.2, 1, 1, None, None, 0, None) == _lib.ERR_EMPTY
    assert lib.gssd_match(1, 100, 1, 1, 2, 900, 500, 0.5, 0.1, 0.2, 1, 1, None, None, 0, None) == _lib.ERR_LIMIT
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_VALUE)
    with pytest.raises(IndexError):
        _lib.check(_lib.ERR_EMPTY)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.ERR_WS)


def test_source_block_and_pipeline_argument_errors():
    """the tcgen05 conv, the PM helpers, the exchange and the pipeline validate their arguments on the host"""
    import ctypes as C
    lib = _lib.load()
    d = _lib.ConvDesc()
    assert lib.gssd_conv_igemm(C.byref(d), None) == _lib.ERR_ARG                       # null tensors
    d.x, d.w, d.y = 1, 1, 1
    d.n_img, d.height, d.width, d.c_in, d.c_out, d.groups, d.taps = 1, 4, 4, 96, 128, 1, 9
    assert lib.gssd_conv_igemm(C.byref(d), None) == _lib.ERR_LIMIT                     # c_in/groups must be a multiple of 64
    d.c_in, d.taps = 128, 5
    assert lib.gssd_conv_igemm(C.byref(d), None) == _lib.ERR_ARG                       # only 1x1 and 3x3
    d.taps, d.c_out = 9, 100
    assert lib.gssd_conv_igemm(C.byref(d), None) == _lib.ERR_LIMIT                     # c_out/groups must be a multiple of 64
    d.c_out, d.width = 128, 3000
    assert lib.gssd_conv_igemm(C.byref(d), None) == _lib.ERR_LIMIT                     # slab of a 3000-wide map does not fit
    d.width, d.y, d.loc, d.conf, d.n_anchor, d.n_cls, d.n_priors, d.c_out = 4, None, 1, 1, 4, 2, 64, 20
    assert lib.gssd_conv_igemm(C.byref(d), None) == _lib.ERR_ARG                       # head: c_out must be anchors * (4 + classes)
    assert lib.gssd_conv_pack_weights(1, 64, 64, 1, 4, None, 1, None) == _lib.ERR_ARG
    assert lib.gssd_nchw_to_pm(None, 1, 64, 4, 4, None, None) == _lib.ERR_ARG
    oh, ow = C.c_int(), C.c_int()
    assert lib.gssd_maxpool_pm(None, 1, 64, 75, 75, 2, 2, 0, 1, None, C.byref(oh), C.byref(ow), None) == 0 and (oh.value, ow.value) == (38, 38)
    assert lib.gssd_maxpool_pm(None, 1, 64, 19, 19, 3, 1, 1, 0, None, C.byref(oh), C.byref(ow), None) == 0 and (oh.value, ow.value) == (19, 19)
    assert lib.gssd_maxpool_pm(None, 1, 60, 19, 19, 3, 1, 1, 0, None, None, None, None) == _lib.ERR_LIMIT
    cfg = _lib.PipeCfg()
    assert lib.gssd_pipe_arena_bytes(C.byref(cfg)) == 0
    cfg.B, cfg.P, cfg.C, cfg.top_k, cfg.max_gt_rows, cfg.depth = 2, 158, 2, 200, 64, 3
    cfg.var0, cfg.var1, cfg.nms_thresh = 0.1, 0.2, 0.45
    assert lib.gssd_pipe_arena_bytes(C.byref(cfg)) > 2 * 158 * (16 + 8 + 8 + 16 + 8) * 3
    h = C.c_void_p()
    assert lib.gssd_pipe_create(C.byref(h), C.byref(cfg), None, None, 0) == _lib.ERR_ARG
    cfg.nms_thresh = 0.0
    assert lib.gssd_pipe_create(C.byref(h), C.byref(cfg), 1, 1, 1 << 30) == _lib.ERR_VALUE   # detection_pytorch_ver_1point5.py:39-40
    x = _lib.Xchg()
    assert lib.gssd_mbox_match_x(1, 100, None, 2, 1, 1, 2, 4, 2, 0.5, 1, 1, C.byref(x), None) == _lib.ERR_ARG   # world == 0
    assert lib.gssd_detect_logits(1, 1, None, 1, 1, 10, 1, 200, 0.1, 0.45, 0.1, 0.2, 1, None, None, None) == _lib.ERR_ARG   # softmax needs C >= 2


def test_layers_module_tree_mirrors_the_reference():
    from grouped_ssd_pytorch_b200 import layers
    from grouped_ssd_pytorch_b200.layers import Detect, L2Norm, MultiBoxLoss, PriorBox, box_utils
    from grouped_ssd_pytorch_b200.layers.functions import Detect as D2, PriorBox as P2
    from grouped_ssd_pytorch_b200.layers.modules import L2Norm as L2, MultiBoxLoss as M2
    assert (D2, P2, L2, M2) == (Detect, PriorBox, L2Norm, MultiBoxLoss)
    for fn in ("point_form", "center_size", "intersect", "jaccard", "match", "encode", "decode", "log_sum_exp", "nms"):
        assert callable(getattr(box_utils, fn))
    # signatures of the reference (box_utils.py:70,174; multibox_loss.py:31-33)
    assert list(inspect.signature(box_utils.match).parameters) == [
        "threshold", "truths", "priors", "variances", "labels", "loc_t", "conf_t", "idx"]
    assert list(inspect.signature(box_utils.nms).parameters) == ["boxes", "scores", "overlap", "top_k"]
    assert list(inspect.signature(MultiBoxLoss.__init__).parameters)[1:] == [
        "num_classes", "overlap_thresh", "prior_for_matching", "bkg_label", "neg_mining", "neg_pos",
        "neg_overlap", "encode_target", "use_gpu"]
    crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)      # train_lesion_multiphase_v2.py:639
    assert crit.variance == [0.1, 0.2] and crit.negpos_ratio == 3 and crit.threshold == 0.5
    assert layers.__name__.endswith("layers")


def test_install_as_layers_resolves_reference_imports():
    saved = {k: sys.modules.get(k) for k in ("layers", "layers.box_utils", "layers.functions", "layers.modules", "data")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        pkg.install_as_layers()
        ns = {}
        exec("from layers import *\nfrom layers.modules import MultiBoxLoss\nfrom layers.box_utils import match, nms\n"
             "from data import v2", ns)
        assert ns["v2"]["name"] == "v2" and "Detect" in ns and "PriorBox" in ns and "L2Norm" in ns
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


REF = "/root/reference/ssd_liverdet"

_DROPIN_SCRIPT = r"""
import os, sys, types
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
dcn = types.ModuleType("dcn_v2"); dcn._DCNv2 = type("_DCNv2", (), {"apply": staticmethod(lambda *a: None)}); sys.modules["dcn_v2"] = dcn
mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
if %(preimport)r:
    import layers                                   # the reference's own package is already imported: install must replace it
    assert "grouped_ssd" not in layers.MultiBoxLoss.__module__
import grouped_ssd_pytorch_b200 as gssd
gssd.install_as_layers()
# the serialised priors of the v2 configuration (the kernel's output, bit-identical to the reference's: test_gpu_parity)
from grouped_ssd_pytorch_b200 import config
from grouped_ssd_pytorch_b200.layers import PriorBox
os.environ["GSSD_PRIOR_CACHE"] = %(cache)r
os.makedirs(%(cache)r, exist_ok=True)
np.save(PriorBox(config.v2)._cache_file(), np.load(os.path.join(%(root)r, "tests", "golden", "priors.npz"))["v2"])
# the imports of models/ssd_multiphase_custom_group.py:5-10 and train_lesion_multiphase_v2.py:14,17
from layers import *
from layers import self_attn
from layers.dcn_v2_custom import DCN
from layers.modules import MultiBoxLoss
from layers.box_utils import match, log_sum_exp, decode, nms
from data import DataSplitter, FISHdetectionV2, detection_collate_v2, BaseTransform, v2
import layers, data
ours = "grouped_ssd_pytorch_b200"
assert MultiBoxLoss.__module__.startswith(ours) and Detect.__module__.startswith(ours) and PriorBox.__module__.startswith(ours)
assert L2Norm.__module__.startswith(ours) and match.__module__.startswith(ours)
assert self_attn.__name__.startswith(ours) and DCN.__module__.startswith(ours)          # GSSD++'s DCN / Self_Attn on our kernels (SURVEY f4)
from layers import spectral_norm as _sn
assert _sn.__file__.startswith(%(ref)r)                                                   # the rest of the reference's package stays importable
assert data.__file__.startswith(%(ref)r), "the reference's data package must stay the one that is imported"
from models.ssd_multiphase_custom_group import build_ssd
import torch
net = build_ssd('train', 300, 2, True, 4, 4, 1, True, False, False, 0, 1, False, False, 1)          # GSSD
assert type(net.L2Norm).__module__.startswith(ours) and type(net.priorbox).__module__.startswith(ours)
assert tuple(net.priors.shape) == (8732, 4) and sum(p.numel() for p in net.parameters()) == 8340084
tst = build_ssd('test', 300, 2, True, 4, 4, 1, True, False, False, 0, 1, False, False, 1)
assert type(tst.detect).__module__.startswith(ours)
pp = build_ssd('train', 300, 2, True, 4, 4, 1, True, True, True, 1, 4, True, False, 1)               # GSSD++ (SA + DCN)
assert sum(p.numel() for p in pp.parameters()) == 18488172
assert type(pp.dcn_list[0]).__module__.startswith(ours) and sorted(n for n, _ in pp.dcn_list[0].named_parameters()) == [
    "bias", "conv_offset_mask.bias", "conv_offset_mask.weight", "weight"]                                # the reference's state-dict names
assert type(pp.self_attn_list[0]).__module__.startswith(ours) and type(pp.self_attn_base_list[0]).__module__.startswith(ours)
assert sorted(pp.self_attn_list[0].state_dict()) == sorted(
    ["sigma"] + ["snconv1x1_%%s.%%s" %% (c, n) for c in ("theta", "phi", "g", "attn") for n in ("bias", "weight_orig", "weight_u", "weight_v")])
crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, True)                                      # train_lesion_multiphase_v2.py:639
print("DROPIN-OK")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container)")
@pytest.mark.parametrize("preimport", [False, True])
def test_install_as_layers_builds_the_reference_model(tmp_path, preimport):
    """INTEGRATION.md option A with the reference's OWN code: after install_as_layers() the model file of the reference
    (models/ssd_multiphase_custom_group.py) imports, its `from layers import *` / `from layers import self_attn` /
    `from layers.dcn_v2_custom import DCN` / `from data import v2` resolve, and build_ssd constructs GSSD and GSSD++ with our
    PriorBox / L2Norm / Detect inside (no GPU here: the priors come from the serialised cache, nothing runs)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = _DROPIN_SCRIPT % dict(root=root, ref=REF, cache=str(tmp_path / "priors"), preimport=preimport)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_reference_visible_errors_without_gpu_work():
    from grouped_ssd_pytorch_b200.layers import Detect, PriorBox
    bad = dict(config.v2); bad["variance"] = [0.0, 0.2]
    with pytest.raises(ValueError):
        PriorBox(bad)                                            # prior_box.py:28-30
    with pytest.raises(ValueError):
        Detect(2, 0, 200, 0.01, 0.0)                             # detection.py:19-20
    with pytest.raises(ValueError):
        Detect.apply(2, 0, 200, 0.01, 0.0, torch.zeros(1, 4, 4), torch.zeros(1, 4, 2), torch.zeros(4, 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only box behaviour")
def test_no_cpu_fallback():
    from grouped_ssd_pytorch_b200.layers import MultiBoxLoss, PriorBox, box_utils
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PriorBox(config.v2).forward()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        box_utils.jaccard(torch.rand(2, 4), torch.rand(3, 4))
    crit = MultiBoxLoss(2, 0.5, True, 0, True, 3, 0.5, False, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit((torch.zeros(1, 4, 4), torch.zeros(1, 4, 2), torch.rand(4, 4)), [torch.rand(1, 5)])


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "synthetic.py", os.path.join(dirpath, f)


def test_synthetic_inputs_are_reproducible():
    from grouped_ssd_pytorch_b200 import synthetic as syn
    a = syn.targets(syn.rng(5), 4)
    b = syn.targets(syn.rng(5), 4)
    assert all((x == y).all() for x, y in zip(a, b))
    assert all(1 <= t.shape[0] <= 5 and t.shape[1] == 5 and (t[:, 2] >= t[:, 0]).all() for t in a)
    s = syn.detect_scores(syn.rng(1), 2, 1000, 2, -4.0)
    assert 0.005 < (s[..., 1] > 0.2).mean() < 0.08


def test_run_layers_leaves_cpu_tensors_to_torch():
    """layers/modules/bn_relu.py on a CPU box: nothing is taken, every module runs as itself (no CUDA call, no error)"""
    import torch.nn as nn
    from grouped_ssd_pytorch_b200.layers.modules import bn_relu as BR
    torch.manual_seed(0)
    mods = nn.ModuleList([nn.Conv2d(4, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(inplace=True), nn.MaxPool2d(2, 2, ceil_mode=True),
                          nn.Conv2d(8, 8, 1), nn.BatchNorm2d(8)]).train()
    import copy
    ref = copy.deepcopy(nn.Sequential(*mods))
    x = torch.randn(2, 4, 7, 7)
    assert not BR.takes(x, mods[1])
    assert torch.equal(BR.run_layers(mods, x), ref(x))
    assert torch.equal(mods[1].running_mean, ref[1].running_mean) and int(mods[1].num_batches_tracked) == 1
    x8 = torch.randn(2, 8, 7, 7)
    assert torch.equal(BR.run_layers(mods, x8, 3, 5), ref[4](ref[3](x8)))                 # a slice of the list
    assert BR._pool_geometry(nn.MaxPool2d(2, 2)) == (2, 2, 0) and BR._pool_geometry(nn.MaxPool2d(3, 1, 1)) == (3, 1, 1)
    assert BR._pool_geometry(nn.MaxPool2d((3, 2))) is None and BR._pool_geometry(nn.MaxPool2d(2, dilation=2)) is None
    assert BR._pool_geometry(nn.MaxPool2d(2, return_indices=True)) is None
    with pytest.raises(NotImplementedError):
        BR.bn_relu(x, mods[1])


def test_provide_dcn_v2_registers_this_packages_operator():
    import subprocess
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import grouped_ssd_pytorch_b200 as g\n"
            "m = g.provide_dcn_v2()\n"
            "import dcn_v2\n"
            "from grouped_ssd_pytorch_b200.layers import dcn_v2_custom as D\n"
            "assert dcn_v2 is m and dcn_v2._DCNv2 is D._DCNv2 and dcn_v2.DCN is D.DCN\n"
            "print('OK')\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container)")
def test_install_as_layers_can_keep_the_references_gssdpp_modules():
    import subprocess
    code = ("import sys, types; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "d = types.ModuleType('dcn_v2'); d._DCNv2 = type('_DCNv2', (), {'apply': None}); sys.modules['dcn_v2'] = d\n"
            "import grouped_ssd_pytorch_b200 as g\n"
            "g.install_as_layers(reference_modules=('dcn_v2_custom', 'self_attn'))\n"
            "from layers.dcn_v2_custom import DCN\n"
            "from layers import self_attn\n"
            "from layers.modules import MultiBoxLoss\n"
            "assert sys.modules['layers.dcn_v2_custom'].__file__.startswith(%r) and self_attn.__file__.startswith(%r)\n"
            "assert MultiBoxLoss.__module__.startswith('grouped_ssd_pytorch_b200')\n"
            "print('OK')\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), REF, REF, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]
