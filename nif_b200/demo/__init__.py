"""Bundled demo datasets (mirror of nif/demo): the two travelling waves are regenerated from their
closed form (tutorial/1_simple_1d_wave.ipynb cell 3, tutorial/2_multi_scale_NIF.ipynb cell 3); the
9 MB cylinder-flow AMR dataset is not redistributed."""
import numpy as np

from ..data.point_wise_data import PointWiseData

__all__ = ["TravelingWave", "TravelingWaveHighFreq"]


def _wave(omega):
    x = np.linspace(0, 1, 200, endpoint=False)
    t = np.linspace(0, 100, 10, endpoint=False)
    xx, tt = np.meshgrid(x, t)
    s = xx - 0.2 - (0.12 / 20) * tt
    u = np.exp(-1000 * s**2) * np.sin(omega * s)
    return np.stack([tt.ravel(), xx.ravel(), u.ravel()], 1).astype(np.float32)


class TravelingWave(PointWiseData):
    """(2000, 3) [t, x, u], standard-normalised (nif/demo/traveling_wave.py:19-36)."""

    def __init__(self):
        d = _wave(4.0)
        super().__init__(d[:, [0]], d[:, [1]], d[:, [2]])
        self.data, self.mean, self.std = self.standard_normalize(self.data_raw)


class TravelingWaveHighFreq(PointWiseData):
    """omega = 400, min-max normalised (nif/demo/traveling_wave_high_freq.py:25-40)."""

    def __init__(self):
        d = _wave(400.0)
        super().__init__(d[:, [0]], d[:, [1]], d[:, [2]])
        self.data, self.mean, self.std = self.minmax_normalize(self.data_raw, n_para=self.n_p, n_x=self.n_x, n_target=1)
