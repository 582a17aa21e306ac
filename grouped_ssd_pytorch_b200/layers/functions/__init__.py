from .detection import Detect, collect_detections
from .prior_box import PriorBox

__all__ = ['Detect', 'PriorBox']
