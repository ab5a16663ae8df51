"""piml_b200 -- B200 (sm_100a) implementation of PIML's per-timestep crowd-rollout hot path.

Host code is Python/PyTorch with the reference's call signatures; every kernel lives in lib/libpiml_b200.so (C ABI in
include/piml_b200.h, sources in piml_b200/csrc).  There is no CPU fallback: compute entry points raise without a GPU.
"""
from . import _lib
from .features import Pedestrians, cos_threshold
from .mlapm import MLAPM
from .sfm import SocialForce, calc_acceleration
from . import metrics, models

__all__ = ["Pedestrians", "MLAPM", "calc_acceleration", "SocialForce", "metrics", "models", "cos_threshold", "_lib"]
__version__ = "0.1.0"
