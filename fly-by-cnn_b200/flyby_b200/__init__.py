"""flyby_b200 -- B200-native fly-by feature extraction (drop-in for the
fly_by_features path of DCBIA-OrthoLab/Fly-by-CNN).

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C-ABI of
include/flyby_b200.h); this package is the Python host that keeps the
reference's class / function / CLI names.
"""
from ._cabi import Context, FlyByError  # noqa: F401

__all__ = ["Context", "FlyByError"]
