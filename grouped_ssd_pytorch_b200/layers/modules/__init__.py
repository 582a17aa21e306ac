from .l2norm import L2Norm
from .multibox_loss import MultiBoxLoss
from .source_block import PM, SourceBlock

__all__ = ['L2Norm', 'MultiBoxLoss', 'SourceBlock', 'PM']
