"""bench.py's workload definition (CPU): the batch fills the last wave of the persistent tensor-core kernels, and the
whole-model parity fixture exists for exactly that batch and configuration."""
import os

import numpy as np

import bench
from tests.helpers import GOLDEN, C2_CFG_S, C2_CFG_P


def test_bench_batch_is_a_whole_number_of_tile_pair_waves():
    # nif_tc_fwd_kernel / nif_tc_bwd_data_kernel: one persistent CTA per SM (148 on B200) over pairs of 128-row tiles
    assert bench.BATCH % (148 * 2 * 128) == 0
    assert bench.N_POINTS // bench.BATCH >= 2  # batches rotate through the point set (inputs larger than L2 per step)


def test_whole_model_fixture_matches_the_benchmark_configuration():
    assert bench.CFG_S == C2_CFG_S and bench.CFG_P == C2_CFG_P
    ref = np.load(os.path.join(GOLDEN, "fullbatch", f"c2_fullbatch_model_grad_{bench.BATCH}.npz"))
    assert np.isfinite(float(ref["loss"])) and any(k.startswith("g:") for k in ref.files)


def test_traffic_capture_is_of_the_benchmark_batch():
    import json
    t = json.load(open(os.path.join(bench.ROOT, "profiles", "r02x_ncu_kernels.json")))
    assert int(t["_batch"]) == bench.BATCH and "nif_tc_bwd_weight_kernel" in t
