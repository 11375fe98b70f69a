"""contrast-you_b200: B200-native (sm_100a) kernels behind Contrast-You's self-supervised loss hot path.

Drop-in modules with the reference's signatures (``contrastyou/losses/contrastive.py``,
``contrastyou/losses/discreteMI.py``, ``contrastyou/projectors``):

    from contrast_you_b200.losses.contrastive import SupConLoss1, SelfPacedSupConLoss
    from contrast_you_b200.losses.discreteMI import IIDSegmentationLoss, IIDLoss
    from contrast_you_b200.projectors import ProjectionHead, DenseProjectionHead, ClusterHead, DenseClusterHead

All loss arithmetic runs in ``libcontrastyou_b200.so`` (C ABI in ``include/contrastyou_b200.h``), loaded with
ctypes by :mod:`._lib`.  There is no CPU or eager fallback: calling a loss without the library, or with CPU
tensors, raises.
"""
from . import _lib  # noqa: F401  (defines load(); does not touch CUDA at import time)

__all__ = ["_lib"]
__version__ = "0.1.0"
