"""ctypes binding of libdrb.so (include/drb.h).  The product path is this library;
there is NO fallback: if it is missing or a call fails, we raise."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdrb.so")

_lib = None

P = c_void_p
_SIGNATURES = {
    "drb_version": ([], c_int),
    "drb_status_string": ([c_int], c_char_p),
    "drb_last_error": ([], c_int),
    "drb_device_sm_count": ([], c_int),
    "drb_sample": ([P, P, c_uint64, c_uint64, P, c_float, c_int, c_int, c_int, c_int, P, P, P, P, P], c_int),
    "drb_sample_sets": ([P, c_uint64, c_uint64, P, c_int, c_int, c_int, c_int, P, P], c_int),
    "drb_sample_backward": ([P, P, c_uint64, c_uint64, P, c_float, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P],
                            c_int),
    "drb_solve_e5": ([P, P, c_int, c_int, c_int, P, P, P, P, P, P], c_int),
    "drb_solve_e5_backward": ([P, P, c_int, c_int, c_int, P, P, P, P, P], c_int),
    "drb_solve_e5_select": ([P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P], c_int),
    "drb_solve_e5_backward_chosen": ([P, P, c_int, c_int, c_int, P, P, P, P, P], c_int),
    "drb_select_closest": ([P, P, P, c_int, c_int, c_int, c_int, P, P, P], c_int),
    "drb_solve_f8": ([P, P, c_int, c_int, c_int, P, P, P], c_int),
    "drb_solve_f8_backward": ([P, P, c_int, c_int, c_int, P, P, P, P], c_int),
    "drb_solve_f7": ([P, P, c_int, c_int, c_int, P, P, P], c_int),
    "drb_refit_e5": ([P, P, P, c_int, c_int, P, P, P], c_int),
    "drb_refit_f8": ([P, P, P, c_int, c_int, P, P, P], c_int),
    "drb_adaptive_select": ([P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double,
                             c_double, P, P, P, P, P], c_int),
    "drb_pose_loss": ([P, P, P, P, P, c_int, c_int, c_int, c_float, P, P, P, P], c_int),
    "drb_recover_pose": ([P, P, P, P, P, c_int, c_int, c_int, c_float, P, P, P, P, P, P, P], c_int),
    "drb_solve_rigid3": ([P, P, c_int, c_int, c_int, c_int, P, P, P], c_int),
    "drb_solve_rigid3_backward": ([P, P, c_int, c_int, c_int, c_int, P, P, P, P], c_int),
    "drb_score_msac": ([P, P, P, P, P, c_int, c_int, c_int, P, P, P], c_int),
    "drb_score_msac_workspace_bytes": ([c_int, c_int, c_int], ctypes.c_size_t),
    "drb_score_msac_workspace_zeroed_bytes": ([c_int, c_int], ctypes.c_size_t),
    "drb_score_msac_stream": ([P, P, P, P, P, c_int, c_int, c_int, P, P, P, ctypes.c_size_t, P], c_int),
    "drb_score_msac_tc_workspace_bytes": ([c_int, c_int], ctypes.c_size_t),
    "drb_score_msac_tc": ([P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, ctypes.c_size_t, P], c_int),
    "drb_copy_h2d_async": ([P, P, ctypes.c_size_t, P], c_int),
    "drb_solve_e5_f64": ([P, P, c_int, c_int, c_int, P, P, P], c_int),
    "drb_score_msac_f64": ([P, P, P, P, c_int, c_int, c_int, c_int, P, P], c_int),
    "drb_best_finalize_f64": ([P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P], c_int),
    "drb_best_finalize": ([P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, P], c_int),
    "drb_episym_forward": ([P, P, P, P, c_int, c_int, c_int, P, P], c_int),
    "drb_episym_backward": ([P, P, P, P, P, c_int, c_int, c_int, P, P], c_int),
    "drb_episym_forward_backward": ([P, P, P, P, P, c_int, c_int, c_int, P, P, P], c_int),
    "drb_rigid_residual_forward": ([P, P, c_int, c_int, c_int, c_float, P, P, P], c_int),
    "drb_rigid_residual_backward": ([P, P, P, c_int, c_int, c_int, P, P], c_int),
    "drb_rigid_residual_forward_backward": ([P, P, P, c_int, c_int, c_int, P, P, P], c_int),
    "drb_gather_backward": ([P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P], c_int),
}

EXPORTS = tuple(_SIGNATURES)


class DrbError(RuntimeError):
    pass


def load():
    """dlopen libdrb.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrbError(
            f"{LIB_PATH} not found: build it with `python -m differentiable_ransac_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        lib = load()
        msg = lib.drb_status_string(status).decode()
        extra = ""
        if status == -4:
            extra = f" (cudaGetLastError={lib.drb_last_error()})"
        raise DrbError(f"{what}: {msg}{extra}")
