"""B200-native GSSD multibox head: the reference's `layers` API over hand-written sm_100a kernels.

    from grouped_ssd_pytorch_b200.layers import PriorBox, Detect, MultiBoxLoss, L2Norm, box_utils

`install_as_layers()` registers the package as top-level `layers` (and `data` if the reference's own is not importable)
so that the reference's `from layers import *` (models/ssd_multiphase_custom_group.py:5) resolves to it.
"""
__version__ = "0.2.0"


def _reference_layers_dir(own_dir, reference_root=None):
    """the reference's own `layers` directory (ssd_liverdet/layers), if it can be found: it still provides the modules
    that are not part of the multibox hot path (self_attn, spectral_norm, dcn_v2_custom)."""
    import os
    import sys
    roots = [reference_root] if reference_root else list(sys.path)
    for root in roots:
        cand = os.path.join(root or ".", "layers")
        if os.path.isfile(os.path.join(cand, "__init__.py")) and not os.path.samefile(cand, own_dir):
            return os.path.abspath(cand)
    return None


def install_as_layers(provide_data=True, reference_root=None, reference_modules=()):
    """Make `layers` (layers/__init__.py:1-2 of the reference) resolve to this package for code imported AFTER the call:
    `layers`, `layers.box_utils`, `layers.functions[.prior_box|.detection|.detection_pytorch_ver_1point5]`,
    `layers.modules[.l2norm|.multibox_loss]`, `layers.dcn_v2_custom` (GSSD++'s deformable convolution on this library's
    kernels instead of the un-vendored `dcn_v2` extension) and `layers.self_attn` (its attention core on this library's
    kernels) become ours; every other submodule of the reference's package — `layers.spectral_norm` — stays importable from
    the reference tree, which is appended to the package search path when it is found on sys.path (or under `reference_root`,
    the directory that holds the reference's `layers/`).  `data` is provided (as our prior-box config module) only when no
    `data` package is importable at all, so that the reference's `from data import DataSplitter, ...`
    (train_lesion_multiphase_v2.py:14) keeps working.  `reference_modules` names submodules (e.g. "dcn_v2_custom") that
    should stay the REFERENCE's even though this package has its own — for A/B measurements against the reference's modules."""
    import importlib.util
    import os
    import sys
    from . import layers as _layers
    from .layers import dcn_v2_custom as _dcn, self_attn as _sa  # noqa: F401  (imported here so that they are aliased below)
    own_dir = os.path.dirname(os.path.abspath(_layers.__file__))
    ref_dir = _reference_layers_dir(own_dir, reference_root)
    if ref_dir is not None and ref_dir not in list(_layers.__path__):
        _layers.__path__.append(ref_dir)
    # drop the reference's hot-path modules if they were imported before us, then alias ours under the top-level names
    prefix = __name__ + ".layers"
    for name in [n for n in sys.modules if n == "layers" or n.startswith("layers.")]:
        mod = sys.modules[name]
        f = getattr(mod, "__file__", None) or ""
        if ref_dir is None or not os.path.abspath(f).startswith(ref_dir) or name in (
                "layers", "layers.box_utils", "layers.functions", "layers.modules", "layers.functions.prior_box",
                "layers.functions.detection", "layers.functions.detection_pytorch_ver_1point5", "layers.modules.l2norm",
                "layers.modules.multibox_loss", "layers.dcn_v2_custom", "layers.self_attn"):
            del sys.modules[name]
    sys.modules["layers"] = _layers
    for name, mod in list(sys.modules.items()):
        if name.startswith(prefix + "."):
            sys.modules["layers" + name[len(prefix):]] = mod
    sys.modules["layers.functions.detection_pytorch_ver_1point5"] = sys.modules[prefix + ".functions.detection"]
    for sub in reference_modules:
        path = os.path.join(ref_dir or "", sub + ".py")
        if ref_dir is None or not os.path.isfile(path):
            raise ImportError("install_as_layers: the reference's layers/%s.py was not found" % sub)
        spec = importlib.util.spec_from_file_location("layers." + sub, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["layers." + sub] = mod
        spec.loader.exec_module(mod)
        setattr(_layers, sub, mod)          # `from layers import self_attn` reads the package attribute, not sys.modules
    if provide_data and "data" not in sys.modules:
        try:
            found = importlib.util.find_spec("data") is not None
        except (ImportError, ValueError):
            found = False
        if not found:
            from . import config as _config
            sys.modules["data"] = _config
    return _layers


def provide_dcn_v2():
    """Register a top-level `dcn_v2` module (the third-party extension dcn_v2_custom.py:13 and utils/try_dcnv2.py:2 import,
    absent from the reference tree) whose `_DCNv2`, `dcn_v2_conv`, `DCNv2` and `DCN` are this package's — for code that imports
    the extension directly.  Does nothing when a real `dcn_v2` is importable."""
    import importlib.util
    import sys
    import types
    if "dcn_v2" in sys.modules:
        return sys.modules["dcn_v2"]
    try:
        if importlib.util.find_spec("dcn_v2") is not None:
            return None
    except (ImportError, ValueError):
        pass
    from .layers import dcn_v2_custom as _dcn
    mod = types.ModuleType("dcn_v2")
    mod._DCNv2, mod.dcn_v2_conv, mod.DCNv2, mod.DCN = _dcn._DCNv2, _dcn.dcn_v2_conv, _dcn.DCNv2, _dcn.DCN
    sys.modules["dcn_v2"] = mod
    return mod
