"""PointWiseData: container + normalisers (reference: nif/data/point_wise_data.py:1-114)."""
import numpy as np


class PointWiseData(object):
    """Rows are [parameters | coordinates | targets | (area weight)]."""

    def __init__(self, parameter_data, x_data, u_data, sample_weight=None):
        cols = [parameter_data, x_data, u_data] + ([sample_weight] if sample_weight is not None else [])
        self.data_raw = np.hstack(cols)
        self.data = None
        self.sample_weight = None
        self.n_p = parameter_data.shape[-1]
        self.n_x = x_data.shape[-1]
        self.n_o = u_data.shape[-1]

    @property
    def parameter(self):
        return self.data[:, : self.n_p]

    @property
    def x(self):
        return self.data[:, self.n_p: self.n_p + self.n_x]

    @property
    def u(self):
        return self.data[:, self.n_p + self.n_x: self.n_p + self.n_x + self.n_o]

    @staticmethod
    def _finish(raw_data, mean, std, area_weighted):
        if area_weighted:  # the last column is a cell area: divide by its mean, do not centre it
            mean[-1] = 0.0
            std[-1] = np.mean(raw_data[:, -1])
            nd = (raw_data - mean) / std
            return nd[:, :-1], mean, std, nd[:, -1]
        return (raw_data - mean) / std, mean, std

    @staticmethod
    def standard_normalize(raw_data, area_weighted=False):
        """zero mean / unit variance per column (point_wise_data.py:51-78)."""
        return PointWiseData._finish(raw_data, raw_data.mean(axis=0), raw_data.std(axis=0), area_weighted)

    @staticmethod
    def minmax_normalize(raw_data, n_para, n_x, n_target, area_weighted=False):
        """inputs -> [-1, 1], targets / max|u| (point_wise_data.py:81-114)."""
        mean, std = raw_data.mean(axis=0), raw_data.std(axis=0)
        lo, hi = raw_data.min(axis=0), raw_data.max(axis=0)
        k = n_para + n_x
        mean[:k] = 0.5 * (lo[:k] + hi[:k])
        std[:k] = 0.5 * (hi[:k] - lo[:k])
        std[k:k + n_target] = np.abs(raw_data[:, k:k + n_target]).max(axis=0)
        return PointWiseData._finish(raw_data, mean, std, area_weighted)
