"""Drop-in for the reference's `ssd_liverdet/layers` package (layers/__init__.py:1-2):
`from layers import *` yields Detect, PriorBox, L2Norm, MultiBoxLoss."""
from . import box_utils  # noqa: F401
from .functions import *  # noqa: F401,F403
from .modules import *  # noqa: F401,F403
from .functions import Detect, PriorBox  # noqa: F401
from .modules import L2Norm, MultiBoxLoss  # noqa: F401
