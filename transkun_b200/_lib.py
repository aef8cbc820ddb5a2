"""ctypes binding of libtranskun_b200.so (include/transkun_b200.h).

There is no CPU or PyTorch fallback: if the library is missing this module
raises, and every entry point raises on a non-zero status.
"""
from __future__ import annotations

import ctypes
import os

# TKB_LIBRARY selects an alternative build of the SAME sources (e.g. the -DTKB_TIMELINE diagnostics build)
_LIB_PATH = os.environ.get("TKB_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc",
                                                          "libtranskun_b200.so")
_lib = None

BACKWARD, FORWARD = 0, 1
SWEEP_VITERBI, SWEEP_LOGSUM = 1, 2

# every symbol include/transkun_b200.h declares
EXPORTS = (
    "tkb_version", "tkb_last_error", "tkb_device_check", "tkb_sweep_workspace_bytes", "tkb_semicrf_sweep",
    "tkb_semicrf_sweep_pitched",
    "tkb_sweep_status", "tkb_semicrf_backtrack", "tkb_semicrf_backtrack_strided", "tkb_semicrf_backtrack_push", "tkb_wait_flags", "tkb_semicrf_marginals", "tkb_semicrf_evalpath",
    "tkb_semicrf_evalpath_grad", "tkb_sip_score", "tkb_sip_score_pitched", "tkb_sip_score_scaled", "tkb_sip_split3", "tkb_sip_backward_prep", "tkb_logmel_workspace_bytes", "tkb_logmel",
    "tkb_upload_lower_triangle",
)


class TkbError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise TkbError(
            f"{_LIB_PATH} is missing: build it with `python -m transkun_b200.build` "
            "(nvcc, sm_100a). transkun_b200 has no CPU or PyTorch fallback.")
    import torch  # noqa: F401  (loads libcudart.so.12 into the process before our library needs it)
    # libcufft.so.11 (frontend): make sure it is resolvable whatever LD_LIBRARY_PATH says
    for cand in ("libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11",
                 os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cufft", "lib", "libcufft.so.11")):
        try:
            ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
            break
        except OSError:
            continue
    L = ctypes.CDLL(_LIB_PATH)
    vp, i, u32, sz, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_float
    L.tkb_version.restype = i
    L.tkb_last_error.restype = ctypes.c_char_p
    L.tkb_device_check.restype = i
    L.tkb_sweep_workspace_bytes.restype = sz
    L.tkb_sweep_workspace_bytes.argtypes = [i, i]
    L.tkb_semicrf_sweep.restype = i
    L.tkb_semicrf_sweep.argtypes = [vp, vp, i, i, i, i, vp, u32, vp, vp, vp, vp]
    L.tkb_semicrf_sweep_pitched.restype = i
    L.tkb_semicrf_sweep_pitched.argtypes = [vp, ctypes.c_int64, vp, i, i, i, i, vp, u32, vp, vp, vp, vp]
    L.tkb_sweep_status.restype = i
    L.tkb_sweep_status.argtypes = [vp, ctypes.POINTER(ctypes.c_int), vp]
    L.tkb_semicrf_backtrack.restype = i
    L.tkb_semicrf_backtrack.argtypes = [vp, i, i, vp, i, vp, vp, vp]
    L.tkb_semicrf_backtrack_strided.restype = i
    L.tkb_semicrf_backtrack_strided.argtypes = [vp, i, i, vp, i, vp, ctypes.c_int64, vp, ctypes.c_int64, vp]
    L.tkb_semicrf_backtrack_push.restype = i
    L.tkb_semicrf_backtrack_push.argtypes = [vp, i, i, vp, i, vp, ctypes.POINTER(vp), ctypes.POINTER(vp), i, i,
                                             ctypes.c_int64, u32, vp, vp, vp]
    L.tkb_wait_flags.restype = i
    L.tkb_wait_flags.argtypes = [vp, i, u32, vp, vp]
    L.tkb_semicrf_marginals.restype = i
    L.tkb_semicrf_marginals.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, vp]
    L.tkb_semicrf_evalpath.restype = i
    L.tkb_semicrf_evalpath.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp]
    L.tkb_semicrf_evalpath_grad.restype = i
    L.tkb_semicrf_evalpath_grad.argtypes = [i, i, vp, vp, vp, f, vp, vp, vp]
    L.tkb_sip_score_pitched.restype = i
    L.tkb_sip_score_pitched.argtypes = [vp, vp, vp, i, i, i, vp, ctypes.c_int64, vp]
    L.tkb_sip_score_scaled.restype = i
    L.tkb_sip_score_scaled.argtypes = [vp, vp, vp, i, i, i, f, vp, ctypes.c_int64, vp]
    L.tkb_sip_split3.restype = i
    L.tkb_sip_split3.argtypes = [vp, vp, ctypes.c_longlong, i, vp, vp, vp]
    L.tkb_sip_backward_prep.restype = i
    L.tkb_sip_backward_prep.argtypes = [vp, ctypes.c_int64, i, i, f, vp, vp, vp]
    L.tkb_sip_score.restype = i
    L.tkb_sip_score.argtypes = [vp, vp, vp, i, i, i, vp, vp]
    i64 = ctypes.c_int64
    L.tkb_logmel_workspace_bytes.restype = sz
    L.tkb_logmel_workspace_bytes.argtypes = [i, i, i, i, i]
    L.tkb_logmel.restype = i
    L.tkb_logmel.argtypes = [vp, i64, i64, i64, i, i, i, i, vp, i, vp, vp, vp, i, i, f, vp, vp, vp]
    L.tkb_upload_lower_triangle.restype = i
    L.tkb_upload_lower_triangle.argtypes = [vp, vp, i, i, i, vp]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tkb_last_error().decode("utf-8", "replace")
        raise TkbError(f"{what} failed with status {rc}: {msg}")
