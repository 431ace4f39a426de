from .deepfm import DeepFM
from .xdeepfm import xDeepFM
from .dcn import DCN
from .autoint import AutoInt
from .sequence import DIN

__all__ = ["DeepFM", "xDeepFM", "DCN", "AutoInt", "DIN"]
