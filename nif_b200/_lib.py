"""ctypes binding of libnif_b200.so (the C ABI in include/nif_b200.h).

There is no fallback: if the library is missing, or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# NIF_B200_LIB: an instrumented build of the same library (tools/tc_trace.py); never a different implementation
LIB_PATH = os.path.abspath(os.environ["NIF_B200_LIB"]) if os.environ.get("NIF_B200_LIB") else os.path.join(_HERE, "libnif_b200.so")
CSRC = os.path.join(_HERE, "csrc")

# every symbol include/nif_b200.h declares
SYMBOLS = (
    "nif_last_error", "nif_version", "nif_query_sizes", "nif_pack", "nif_forward", "nif_forward_tangent", "nif_forward_tangent2",
    "nif_forward_given_w", "nif_mse_backward", "nif_mse_backward_ev", "nif_backward", "nif_adam_step", "nif_adam_step_dev", "nif_measure_fp32_peak",
    "nif_trunk_query", "nif_trunk_forward", "nif_trunk_backward", "nif_trunk_kernel_path",
    "nif_trunk_ew_forward", "nif_trunk_ew_ws_floats", "nif_trunk_ew_backward",
    "nif_sobolev_query", "nif_sobolev_query_dirs", "nif_forward_tangent_save", "nif_sobolev_backward", "nif_sobolev_backward_dirs",
    "nif_crc32c",
    "nif_profile_begin", "nif_profile_end", "nif_adabelief_step", "nif_lion_step", "nif_centralize_gradient", "nif_adam_step_multimem",
)

NIF_MAX_DIR = 4  # csrc/nif_common.cuh
VARIANT = {"nif": 0, "siren": 1, "siren_res": 2}
ACT = {None: 0, "linear": 0, "sine": 1, "swish": 2, "tanh": 3, "relu": 4, "sigmoid": 5}


class NifError(RuntimeError):
    pass


class Desc(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("act", C.c_int32), ("si", C.c_int32), ("so", C.c_int32), ("n", C.c_int32),
        ("l", C.c_int32), ("K", C.c_int32), ("omega0", C.c_float), ("dtype_compute", C.c_int32),
        ("acc_rows", C.c_int32),
    ]


class TrunkDesc(C.Structure):
    _fields_ = [("pi", C.c_int32), ("latent", C.c_int32), ("units", C.c_int32), ("nlayers", C.c_int32),
                ("act", C.c_int32)]


class Sizes(C.Structure):
    _fields_ = [
        ("po_dim", C.c_int64), ("n_layers", C.c_int64), ("np", C.c_int64), ("packed_floats", C.c_int64),
        ("save_floats_per_row", C.c_int64), ("grad_ws_floats", C.c_int64), ("tile_rows", C.c_int64), ("kernel_path", C.c_int64),
    ]


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into nif_b200/libnif_b200.so (in-tree)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise NifError("building libnif_b200.so failed (see output above)")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NifError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C nif_b200/csrc`). "
            "nif_b200 has no CPU or PyTorch fallback."
        )
    L = C.CDLL(LIB_PATH)
    P, F, I32, I64, VP = C.POINTER(C.c_float), C.c_float, C.c_int32, C.c_int64, C.c_void_p
    DP = C.POINTER(Desc)
    L.nif_last_error.restype = C.c_char_p
    L.nif_last_error.argtypes = []
    L.nif_version.restype = C.c_int
    L.nif_query_sizes.argtypes = [DP, I64, C.POINTER(Sizes)]
    L.nif_pack.argtypes = [DP, I64, VP, VP, VP, VP]
    L.nif_forward.argtypes = [DP, I64, I64, VP, VP, I32, VP, VP, VP, VP]
    L.nif_forward_tangent.argtypes = [DP, I64, VP, VP, VP, I32, VP, VP, VP, VP, VP]
    L.nif_forward_tangent2.argtypes = [DP, I64, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP]
    L.nif_forward_given_w.argtypes = [DP, I64, VP, VP, VP, VP]
    L.nif_mse_backward.argtypes = [DP, I64, VP, VP, VP, VP, VP, VP, VP, F, VP, VP, VP, F, VP, VP, VP]
    L.nif_mse_backward_ev.argtypes = [DP, I64, VP, VP, VP, VP, VP, VP, VP, F, VP, VP, VP, F, VP, VP, VP, VP]
    L.nif_backward.argtypes = [DP, I64, VP, VP, VP, VP, VP, VP, VP, F, VP, VP, VP]
    L.nif_adam_step.argtypes = [I64, VP, VP, VP, VP, C.c_double, C.c_double, C.c_double, C.c_double, I64, F, F, F, VP]
    L.nif_adam_step_dev.argtypes = [I64, VP, VP, VP, VP, VP, C.c_double, C.c_double, C.c_double, F, F, F, VP]
    L.nif_measure_fp32_peak.argtypes = [C.POINTER(C.c_double)]
    L.nif_sobolev_query.argtypes = [DP, I64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.nif_forward_tangent_save.argtypes = [DP, I64, VP, VP, VP, I32, VP, VP, VP, VP, VP, VP]
    L.nif_sobolev_backward.argtypes = [DP, I64, VP, VP, VP, VP, VP, VP, VP, VP, VP, F, VP, VP, VP]
    L.nif_sobolev_query_dirs.argtypes = [DP, I64, I32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.nif_sobolev_backward_dirs.argtypes = [DP, I64, VP, VP, I32, VP, C.c_uint32, VP, VP, VP, VP, VP, VP, VP, F, VP, VP, VP, VP]
    TP, I64P = C.POINTER(TrunkDesc), C.POINTER(C.c_int64)
    L.nif_trunk_query.argtypes = [TP, I64, I64P, I64P, I64P, I64P]
    L.nif_trunk_kernel_path.argtypes = [TP]
    L.nif_trunk_forward.argtypes = [TP, I64, VP, VP, VP, VP, VP, VP]
    L.nif_trunk_backward.argtypes = [TP, I64, VP, VP, VP, VP, VP, F, VP, VP, VP]
    L.nif_trunk_ew_forward.argtypes = [I64, I32, I32, VP, VP, VP, VP, VP, VP]
    L.nif_trunk_ew_ws_floats.argtypes = [I32]
    L.nif_trunk_ew_backward.argtypes = [I64, I32, I32, VP, VP, VP, VP, VP, VP, VP, VP, VP]
    for name in SYMBOLS:
        getattr(L, name)  # raises AttributeError if the build is stale
        if name not in ("nif_last_error", "nif_crc32c"):
            getattr(L, name).restype = C.c_int
    D = C.c_double
    L.nif_adabelief_step.argtypes = [I64, VP, VP, VP, VP, VP, D, D, D, D, I64, I32, D, D, F, F, F, VP]
    L.nif_lion_step.argtypes = [I64, VP, VP, VP, D, D, D, D, F, F, F, VP]
    L.nif_centralize_gradient.argtypes = [I64, I64, VP, VP]
    L.nif_adam_step_multimem.argtypes = [I64, I32, I32, VP, VP, VP, VP, VP, D, VP, D, D, D, I64, F, F, F, VP]
    L.nif_profile_begin.argtypes = []
    L.nif_profile_end.argtypes = [C.c_char_p, C.c_int64]
    L.nif_crc32c.restype = C.c_uint32
    L.nif_crc32c.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32]
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().nif_last_error().decode("utf-8", "replace")
        raise NifError(f"{what} failed with code {rc}: {msg}")
