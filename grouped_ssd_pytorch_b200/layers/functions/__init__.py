from .detection import Detect
from .evaluation import ap_iobb, collect_detections, evaluate_detections
from .prior_box import PriorBox

__all__ = ['Detect', 'PriorBox']
