"""Host-side data helpers (mirror of nif/data/__init__.py; the TFRecord reader is out of scope)."""
from .point_wise_data import PointWiseData

__all__ = ["PointWiseData"]
