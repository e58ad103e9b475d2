"""bench.py's workload definition (CPU): the batch SURVEY 8(d) fixes, whole-model parity fixtures for it and for the
wave-aligned batch the profiles also report, and ncu traffic captures that name their batch."""
import json
import os

import numpy as np

import bench
from tests.helpers import GOLDEN, C2_CFG_S, C2_CFG_P

WAVE_BATCH = 113664  # 148 SMs x 3 pairs of 128-row tiles (`python bench.py --batch 113664`, profiles/r02x_*, r02y_*)


def test_bench_batch_is_the_survey_batch_and_the_wave_batch_is_whole_waves():
    assert bench.BATCH == 65536  # C2 throughput at B = 65 536
    # nif_tc_fwd_kernel / nif_tc_bwd_data_kernel: one persistent CTA per SM (148 on B200) over pairs of 128-row tiles
    assert WAVE_BATCH % (148 * 2 * 128) == 0
    assert bench.N_POINTS // WAVE_BATCH >= 2  # batches rotate through the point set (inputs larger than L2 per step)


def test_whole_model_fixtures_match_the_benchmark_configuration():
    assert bench.CFG_S == C2_CFG_S and bench.CFG_P == C2_CFG_P
    for fn in ("c2_fullbatch_model_grad.npz", f"c2_fullbatch_model_grad_{WAVE_BATCH}.npz"):
        ref = np.load(os.path.join(GOLDEN, "fullbatch", fn))
        assert np.isfinite(float(ref["loss"])) and any(k.startswith("g:") for k in ref.files)


def test_traffic_captures_name_their_batch():
    t = json.load(open(os.path.join(bench.ROOT, "profiles", "r02x_ncu_kernels.json")))
    assert int(t["_batch"]) == WAVE_BATCH and "nif_tc_bwd_weight_kernel" in t
    t0 = json.load(open(os.path.join(bench.ROOT, "profiles", "r02_ncu_kernels.json")))
    assert "_batch" not in t0 and "nif_tc_bwd_weight_kernel" in t0  # the 65 536-row capture of bench.py's default batch
