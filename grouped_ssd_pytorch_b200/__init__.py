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


def install_as_layers(provide_data=True, reference_root=None):
    """Make `layers` (layers/__init__.py:1-2 of the reference) resolve to this package for code imported AFTER the call:
    `layers`, `layers.box_utils`, `layers.functions[.prior_box|.detection|.detection_pytorch_ver_1point5]`,
    `layers.modules[.l2norm|.multibox_loss]` become ours; every other submodule of the reference's package — `layers.self_attn`,
    `layers.spectral_norm`, `layers.dcn_v2_custom` (models/ssd_multiphase_custom_group.py:6,8) — stays importable from the
    reference tree, which is appended to the package search path when it is found on sys.path (or under `reference_root`,
    the directory that holds the reference's `layers/`).  `data` is provided (as our prior-box config module) only when no
    `data` package is importable at all, so that the reference's `from data import DataSplitter, ...`
    (train_lesion_multiphase_v2.py:14) keeps working."""
    import importlib.util
    import os
    import sys
    from . import layers as _layers
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
                "layers.modules.multibox_loss"):
            del sys.modules[name]
    sys.modules["layers"] = _layers
    for name, mod in list(sys.modules.items()):
        if name.startswith(prefix + "."):
            sys.modules["layers" + name[len(prefix):]] = mod
    sys.modules["layers.functions.detection_pytorch_ver_1point5"] = sys.modules[prefix + ".functions.detection"]
    if provide_data and "data" not in sys.modules:
        try:
            found = importlib.util.find_spec("data") is not None
        except (ImportError, ValueError):
            found = False
        if not found:
            from . import config as _config
            sys.modules["data"] = _config
    return _layers
