"""Golden vectors for PointWiseData (nif/data/point_wise_data.py) produced by importing the UNMODIFIED reference file (it
needs numpy only).  Build container only:   python tests/golden/make_golden_pointwise.py"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_pwd", "/root/reference/nif/data/point_wise_data.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
P = ref.PointWiseData

rng = np.random.default_rng(11)
raw = np.hstack([rng.normal(2.0, 3.0, (200, 2)), rng.uniform(-5, 1, (200, 3)), rng.normal(0, 0.1, (200, 2))])
rawa = np.hstack([raw, rng.uniform(0.5, 2.0, (200, 1))])
out = {"raw": raw, "rawa": rawa}
d, m, s = P.standard_normalize(raw.copy())
out.update(std_data=d, std_mean=m, std_std=s)
d, m, s, w = P.standard_normalize(rawa.copy(), area_weighted=True)
out.update(stda_data=d, stda_mean=m, stda_std=s, stda_w=w)
d, m, s = P.minmax_normalize(raw.copy(), 2, 3, 2)
out.update(mm_data=d, mm_mean=m, mm_std=s)
d, m, s, w = P.minmax_normalize(rawa.copy(), 2, 3, 2, area_weighted=True)
out.update(mma_data=d, mma_mean=m, mma_std=s, mma_w=w)
obj = P(raw[:, :2], raw[:, 2:5], raw[:, 5:7], rawa[:, -1:])
obj.data = obj.data_raw
out.update(obj_parameter=obj.parameter, obj_x=obj.x, obj_u=obj.u, obj_raw=obj.data_raw)
np.savez_compressed(os.path.join(HERE, "pointwise", "pointwise.npz"), **out)
print({k: np.asarray(v).shape for k, v in out.items()})
