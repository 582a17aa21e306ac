"""B200-native GSSD multibox head: the reference's `layers` API over hand-written sm_100a kernels.

    from grouped_ssd_pytorch_b200.layers import PriorBox, Detect, MultiBoxLoss, L2Norm, box_utils

`install_as_layers()` registers the package as top-level `layers` (and `data` if absent) so that the
reference's `from layers import *` (models/ssd_multiphase_custom_group.py:5) resolves to it.
"""
__version__ = "0.1.0"


def install_as_layers(provide_data=True):
    import sys
    from . import layers as _layers
    sys.modules["layers"] = _layers
    for name in ("box_utils", "functions", "modules"):
        sys.modules["layers." + name] = getattr(_layers, name)
    if provide_data and "data" not in sys.modules:
        from . import config as _config
        sys.modules["data"] = _config
    return _layers
