"""nif_b200 -- B200-native engine for the Neural Implicit Flow hot path.

Public surface mirrors `nif` (pswpswpsw/nif, nif/__init__.py:19-28): NIF, NIFMultiScale, data, demo.
Everything numerical runs in libnif_b200.so (hand-written sm_100a CUDA); there is no fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .ops import FusedShapeNet, adam_step, fused_shapenet  # noqa: F401

__all__ = ["FusedShapeNet", "fused_shapenet", "adam_step"]

from .model import NIF, NIFMultiScale, NIFMultiScaleLastLayerParameterized  # noqa: E402,F401
from .keras_like import Adam, Callback, Dataset, LearningRateScheduler, Model, SobolevMSE  # noqa: E402,F401
from .layers import HessianLayer, JacobianLayer  # noqa: E402,F401
from .distributed import DataParallel  # noqa: E402,F401
from . import data, demo, optimizers  # noqa: E402,F401

__all__ += ["NIF", "NIFMultiScale", "NIFMultiScaleLastLayerParameterized", "optimizers", "Adam", "Callback", "Dataset", "LearningRateScheduler", "Model",
            "DataParallel", "data", "demo", "SobolevMSE", "JacobianLayer", "HessianLayer"]
