"""Host-side data helpers (mirror of nif/data/__init__.py)."""
from .point_wise_data import PointWiseData
from .tfr_dataset import TFRDataset

__all__ = ["PointWiseData", "TFRDataset"]
