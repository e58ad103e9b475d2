"""The C-ABI library loads without a GPU and exports every symbol include/nif_b200.h declares;
descriptor validation and size queries are host-only and are checked here."""
import ctypes as C
import os
import re

import pytest

from oracle import nif_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nif_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nif_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from nif_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 10
    assert sorted(_lib.SYMBOLS) == names
    for n in names:
        assert getattr(L, n) is not None


@pytest.mark.parametrize("variant,si,so,n,l,K", [(1, 2, 1, 64, 4, 32), (1, 3, 3, 128, 6, 64), (0, 1, 1, 30, 2, 1),
                                                 (2, 2, 3, 12, 2, 3), (1, 1, 1, 64, 4, 32), (1, 3, 1, 128, 6, 64)])
def test_query_sizes(variant, si, so, n, l, K):
    from nif_b200 import _lib
    L = _lib.lib()
    d = _lib.Desc(variant, 2, si, so, n, l, K, 30.0, 0, 0)
    s = _lib.Sizes()
    assert L.nif_query_sizes(C.byref(d), 4096, C.byref(s)) == 0
    H = 2 * l if variant == 2 else l
    assert s.po_dim == O.po_dim(si, so, n, l, variant == 2)
    assert s.n_layers == H + 2
    assert s.np == (32 if n <= 32 else 64 if n <= 64 else 128)
    assert s.save_floats_per_row == 2 * (H + 1) * s.np
    assert s.packed_floats % 32 == 0
    assert s.packed_floats >= (2 * H * s.np * s.np + si * s.np + s.np * so + (H + 2) * s.np) * (K + 1)
    assert s.grad_ws_floats > (H + 1) * 4096 * s.np


@pytest.mark.parametrize("field,value", [("variant", 7), ("si", 0), ("si", 99), ("so", 0), ("so", 99), ("n", 0),
                                         ("n", 129), ("l", -1), ("K", -1), ("act", 42), ("dtype_compute", 3)])
def test_bad_descriptor_is_rejected(field, value):
    from nif_b200 import _lib
    L = _lib.lib()
    d = _lib.Desc(1, 1, 2, 1, 64, 4, 32, 30.0, 0, 0)
    setattr(d, field, value)
    s = _lib.Sizes()
    rc = L.nif_query_sizes(C.byref(d), 1, C.byref(s))
    assert rc in (-1, -3)
    assert len(L.nif_last_error()) > 0


def test_null_arguments_fail_before_touching_the_device():
    from nif_b200 import _lib
    L = _lib.lib()
    d = _lib.Desc(1, 1, 2, 1, 64, 4, 32, 30.0, 0, 0)
    assert L.nif_pack(C.byref(d), 1, None, None, None, None) == -2
    assert L.nif_forward(C.byref(d), 1, 16, None, None, 0, None, None, None, None) == -2
    assert L.nif_adam_step(16, None, None, None, None, 1e-3, 0.9, 0.999, 1e-7, 1, 0.0, 0.0, 1.0, None) == -2
    assert L.nif_adam_step(16, None, None, None, None, 1e-3, 0.9, 0.999, 1e-7, 0, 0.0, 0.0, 1.0, None) == -2  # t >= 1
    assert L.nif_forward(C.byref(d), 1, 0, None, None, 0, None, None, None, None) == 0  # empty batch is a no-op


def test_null_arguments_of_the_later_entry_points():
    """nif_forward_tangent2 (HessianLayer), nif_adam_step_dev (graph-replayed steps), the Sobolev pair and the trunk:
    same conventions -- -2 with a message before anything is enqueued, 0 for an empty batch."""
    from nif_b200 import _lib
    L = _lib.lib()
    d = _lib.Desc(1, 1, 2, 1, 64, 4, 32, 30.0, 0, 0)
    assert L.nif_forward_tangent2(C.byref(d), 8, None, None, None, None, None, None, None, None, None, None) == -2
    assert b"nif_forward_tangent2" in L.nif_last_error()
    assert L.nif_forward_tangent2(C.byref(d), 0, None, None, None, None, None, None, None, None, None, None) == 0
    assert L.nif_forward_tangent2(C.byref(d), -1, None, None, None, None, None, None, None, None, None, None) == -2
    assert L.nif_adam_step_dev(16, None, None, None, None, None, 0.9, 0.999, 1e-7, 0.0, 0.0, 1.0, None) == -2
    assert L.nif_adam_step_dev(0, None, None, None, None, None, 0.9, 0.999, 1e-7, 0.0, 0.0, 1.0, None) == 0
    assert L.nif_forward_tangent(C.byref(d), 8, None, None, None, 5, None, None, None, None, None) == -2  # n_dir > 4
    assert L.nif_sobolev_backward(C.byref(d), 8, None, None, None, None, None, None, None, None, None, 0.0, None, None,
                                  None) == -2
    # the multi-direction Sobolev pair: direction count in [1, 4], sizes grow with the directions, empty batch is a no-op
    assert L.nif_sobolev_backward_dirs(C.byref(d), 8, None, None, 0, None, 0, None, None, None, None, None, None, None, 0.0,
                                       None, None, None, None) == -2
    assert L.nif_sobolev_backward_dirs(C.byref(d), 8, None, None, 5, None, 0, None, None, None, None, None, None, None, 0.0,
                                       None, None, None, None) == -2
    assert L.nif_sobolev_backward_dirs(C.byref(d), 8, None, None, 2, None, 0, None, None, None, None, None, None, None, 0.0,
                                       None, None, None, None) == -2  # null pointers
    assert L.nif_sobolev_backward_dirs(C.byref(d), 0, None, None, 2, None, 0, None, None, None, None, None, None, None, 0.0,
                                       None, None, None, None) == 0
    s1, s3, w1, w3 = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    assert L.nif_sobolev_query_dirs(C.byref(d), 1024, 1, C.byref(s1), C.byref(w1)) == 0
    assert L.nif_sobolev_query_dirs(C.byref(d), 1024, 3, C.byref(s3), C.byref(w3)) == 0
    assert L.nif_sobolev_query_dirs(C.byref(d), 1024, 9, C.byref(s3), C.byref(w3)) == -2
    assert s1.value == 4 * 5 * 64 and s3.value == 8 * 5 * 64 and w3.value == w1.value  # (2 + 2 n_dir)(H + 1) NP per row
    t = _lib.TrunkDesc(1, 32, 64, 4, 2)
    assert L.nif_trunk_forward(C.byref(t), 8, None, None, None, None, None, None) == -2
    bad = _lib.TrunkDesc(1, 32, 65, 4, 2)  # wider than the fused trunk kernels
    n = C.c_int64(0)
    assert L.nif_trunk_query(C.byref(bad), 8, C.byref(n), None, None, None) == -3
    # a misaligned pointer is refused too
    buf = (C.c_float * 64)()
    addr = C.addressof(buf)
    mis = C.c_void_p(addr + 4 if addr % 16 == 0 else addr + (16 - addr % 16) + 4)
    assert L.nif_adam_step(8, mis, mis, mis, mis, 1e-3, 0.9, 0.999, 1e-7, 1, 0.0, 0.0, 1.0, None) == -2
    assert b"aligned" in L.nif_last_error()


def test_gradient_workspace_holds_the_capped_batch_splits():
    """The tensor-core weight-gradient kernel cuts the batch into splits of at most 4096 rows (bounded TMEM accumulation
    chains); nif_query_sizes must size the workspace for that many partials [H][K+1][NP][NP] on top of the da slots."""
    from nif_b200 import _lib
    L = _lib.lib()
    d = _lib.Desc(1, 1, 2, 1, 64, 4, 32, 30.0, 2, 0)  # the C2 head on the tensor-core path
    prev = 0
    for B in (512, 4096, 65536, 1 << 20):
        s = _lib.Sizes()
        assert L.nif_query_sizes(C.byref(d), B, C.byref(s)) == 0
        H, K1, NP = 4, 33, 64
        da = (H + 1) * ((B + 63) // 64 * 64) * NP
        partials = max(1, -(-B // 4096)) * H * K1 * NP * NP
        assert s.grad_ws_floats >= da + partials, (B, s.grad_ws_floats, da, partials)
        assert s.grad_ws_floats > prev
        prev = s.grad_ws_floats
