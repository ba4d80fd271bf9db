"""ctypes binding of include/gritas_b200.h -- the same entry points the Fortran ISO_C_BINDING
module (fortran/m_gritas_b200.F90) binds.  Arrays may be numpy (host) arrays, torch CUDA
tensors (used in place on the device) or raw integer addresses.

There is no CPU path here: load() raises if libgritas_b200.so is missing, and
gritas_b200_create fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import build as _build

MODE_FAST = 0
MODE_DETERMINISTIC = 1
FLAG_ASYNC = 0x100
FLAG_NO_AGGREGATE = 0x200
FLAG_HOLD_INPUTS = 0x400
FLAG_NO_STAGING = 0x800
FLAG_FORCE_TILED = 0x1000
FLAG_MINMAX = 0x2000
FLAG_NO_WRITEBACK = 0x4000
FLAG_DET_SORT = 0x8000

OK = 0
ERR_LEVEL = 7
ERR_NR = 8
ERR_CUDA = -1
ERR_ARG = -2
ERR_NCCL = -3

IOPT_OMF = 1
IOPT_OMA = 3
IOPT_OMF_OMA = 101

# every symbol include/gritas_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "gritas_b200_create", "gritas_b200_destroy", "gritas_b200_pint", "gritas_b200_kxkt", "gritas_b200_accum",
    "gritas_b200_accum_ods", "gritas_b200_accum_odsx", "gritas_b200_flush", "gritas_b200_barrier", "gritas_b200_sync", "gritas_b200_reset", "gritas_b200_norm", "gritas_b200_norm_3ch", "gritas_b200_presort", "gritas_b200_norm_device",
    "gritas_b200_get_raw", "gritas_b200_get_minmax", "gritas_b200_set_mask", "gritas_b200_add_raw", "gritas_b200_nccl_unique_id", "gritas_b200_comm_init",
    "gritas_b200_allreduce", "gritas_b200_norm_reduce", "gritas_b200_norm_reduce_device", "gritas_b200_host_alloc", "gritas_b200_host_free", "gritas_b200_last_error",
    "gritas_b200_stream", "gritas_b200_set_stream", "gritas_b200_tiled", "gritas_b200_fused_norm", "gritas_b200_launch_count", "gritas_b200_h2d_bytes",
    "gritas_b200_d2h_bytes", "gritas_b200_version",
    "gritas_b200_peer_export", "gritas_b200_peer_attach", "gritas_b200_slab", "gritas_b200_norm_sharded",
    "gritas_b200_norm_sharded_device", "gritas_b200_host_register", "gritas_b200_host_unregister", "gritas_b200_wait",
    "gritas_b200_scores", "gritas_b200_score_regions", "gritas_b200_accum_odsx_sorted", "gritas_b200_set_call_order", "gritas_b200_accum_odsx_meanres", "gritas_b200_export_planes",
    "gritas_b200_xcov_create", "gritas_b200_xcov_destroy", "gritas_b200_xcov_reset", "gritas_b200_xcov_accum",
    "gritas_b200_xcov_norm", "gritas_b200_xcov_get_raw", "gritas_b200_xcov_last_error", "gritas_b200_xcov_launch_count",
    "gritas_b200_xcov_prs_create", "gritas_b200_xcov_prs_accum",
]
ERR_PROFILE = 9
ERR_ORDER = 10
PEER_BLOB_BYTES = 512

OPT_TCH = 1
OPT_SELF = 2


class OdsCols(C.Structure):
    """struct gritas_b200_ods of include/gritas_b200.h"""
    _fields_ = [(k, C.c_void_p) for k in ("lat", "lon", "lev", "obs", "omf", "oma", "xvec", "xm", "time", "kt", "kx", "qcexcl")]


_lib = None
_P = C.c_void_p
_I = C.c_int
_D = C.c_double


def load(build_if_missing: bool = False):
    """Loads libgritas_b200.so.  Fails loudly when it is not there (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RuntimeError(f"{_build.LIB} is missing: run `python -m gritas_b200.build` (nvcc, sm_100a). "
                           "gritas_b200 has no CPU implementation.")
    L = C.CDLL(_build.LIB, mode=C.RTLD_GLOBAL)
    L.gritas_b200_create.argtypes = [C.POINTER(_P)] + [_I] * 7 + [_P, _P, _P, _I, _I, _I, _I]
    L.gritas_b200_destroy.argtypes = [_P]
    L.gritas_b200_destroy.restype = None
    L.gritas_b200_pint.argtypes = [_I, _P, _P, _P]
    L.gritas_b200_kxkt.argtypes = [_I] * 5 + [_P, _P, _P, _I, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)]
    L.gritas_b200_accum.argtypes = [_P, _I] + [_P] * 7 + [C.POINTER(_D)]
    L.gritas_b200_accum_ods.argtypes = [_P, _I] + [_P] * 11 + [_I] * 5 + [_D, _D, _I, _I, _P, C.POINTER(_D)]
    L.gritas_b200_accum_odsx.argtypes = [_P, _I, _P] + [_I] * 6 + [_D, _D, _I, _I, _P, C.POINTER(_D)]
    L.gritas_b200_sync.argtypes = [_P, C.POINTER(_D)]
    L.gritas_b200_wait.argtypes = [_P, C.POINTER(_D)]
    L.gritas_b200_reset.argtypes = [_P]
    L.gritas_b200_flush.argtypes = [_P]
    L.gritas_b200_barrier.argtypes = [_P]
    L.gritas_b200_norm.argtypes = [_P, _I, _I] + [_P] * 10
    L.gritas_b200_norm_3ch.argtypes = [_P, _I, _I] + [_P] * 10
    L.gritas_b200_presort.argtypes = [_P, _I] + [_P] * 8 + [_I, _P]
    L.gritas_b200_norm_device.argtypes = [_P, _I, _I] + [C.POINTER(_P)] * 10
    L.gritas_b200_get_raw.argtypes = [_P] + [_P] * 8
    L.gritas_b200_add_raw.argtypes = [_P] + [_P] * 8
    L.gritas_b200_get_minmax.argtypes = [_P] + [_P] * 6
    L.gritas_b200_set_mask.argtypes = [_P, _P]
    L.gritas_b200_nccl_unique_id.argtypes = [_P]
    L.gritas_b200_comm_init.argtypes = [_P, _I, _I, _P]
    L.gritas_b200_allreduce.argtypes = [_P]
    L.gritas_b200_norm_reduce.argtypes = [_P, _I, _I, _I] + [_P] * 10
    L.gritas_b200_norm_reduce_device.argtypes = [_P, _I, _I, _I, C.POINTER(_P)]
    L.gritas_b200_peer_export.argtypes = [_P, _P]
    L.gritas_b200_peer_attach.argtypes = [_P, _I, _I, _P]
    L.gritas_b200_slab.argtypes = [_P, _I, _I] + [C.POINTER(_I)] * 4
    L.gritas_b200_norm_sharded.argtypes = [_P, _I, _I] + [_P] * 10
    L.gritas_b200_norm_sharded_device.argtypes = [_P, _I, _I, C.POINTER(_P)]
    L.gritas_b200_host_register.argtypes = [_P, C.c_size_t]
    L.gritas_b200_host_unregister.argtypes = [_P]
    L.gritas_b200_accum_odsx_sorted.argtypes = [_P, _I, _P, _P] + [_I] * 6 + [_D, _D, _I, _I, _P, _P, C.POINTER(_D)]
    L.gritas_b200_accum_odsx_meanres.argtypes = [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _D, _D, _I, _I, _P, C.POINTER(_D)]
    L.gritas_b200_export_planes.argtypes = [_P]
    L.gritas_b200_set_call_order.argtypes = [_P, C.c_uint64]
    L.gritas_b200_scores.argtypes = [_P, _I, _P, _I, _I, _P]
    L.gritas_b200_score_regions.argtypes = [_I, _I, _P, _P]
    L.gritas_b200_xcov_create.argtypes = [C.POINTER(_P), _I, _I, _P, _I, _I, _I, _P, _I, _I, _I]
    L.gritas_b200_xcov_prs_create.argtypes = [C.POINTER(_P), _I, _P, _P, _I, _I, _I, _I]
    L.gritas_b200_xcov_prs_accum.argtypes = [_P, _I, _P, _P, _P, C.POINTER(_I), C.POINTER(_D)]
    L.gritas_b200_xcov_destroy.argtypes = [_P]
    L.gritas_b200_xcov_destroy.restype = None
    L.gritas_b200_xcov_reset.argtypes = [_P]
    L.gritas_b200_xcov_accum.argtypes = [_P, _I, _P, _P, _P, _P, _P, C.POINTER(_I)]
    L.gritas_b200_xcov_norm.argtypes = [_P, _I, _I, _I, _P, _P, _P]
    L.gritas_b200_xcov_get_raw.argtypes = [_P, _P, _P, _P]
    L.gritas_b200_xcov_last_error.argtypes = [_P]
    L.gritas_b200_xcov_last_error.restype = C.c_char_p
    L.gritas_b200_xcov_launch_count.argtypes = [_P]
    L.gritas_b200_xcov_launch_count.restype = C.c_uint64
    L.gritas_b200_host_alloc.argtypes = [C.c_size_t]
    L.gritas_b200_host_alloc.restype = _P
    L.gritas_b200_host_free.argtypes = [_P]
    L.gritas_b200_host_free.restype = None
    L.gritas_b200_last_error.argtypes = [_P]
    L.gritas_b200_last_error.restype = C.c_char_p
    L.gritas_b200_stream.argtypes = [_P]
    L.gritas_b200_stream.restype = _P
    L.gritas_b200_set_stream.argtypes = [_P, _P]
    for nm in ("gritas_b200_tiled", "gritas_b200_fused_norm", "gritas_b200_launch_count", "gritas_b200_h2d_bytes", "gritas_b200_d2h_bytes"):
        getattr(L, nm).argtypes = [_P]
        getattr(L, nm).restype = C.c_uint64
    L.gritas_b200_tiled.argtypes = [_P]
    L.gritas_b200_fused_norm.argtypes = [_P]
    L.gritas_b200_fused_norm.restype = C.c_int
    L.gritas_b200_version.argtypes = []
    _lib = L
    return L


def ptr(a, dtype=None):
    """Address of a numpy array / torch tensor / int / None, checking dtype and contiguity."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        if dtype is not None and a.dtype != np.dtype(dtype):
            raise TypeError(f"expected {np.dtype(dtype)}, got {a.dtype}")
        if not (a.flags.f_contiguous or a.flags.c_contiguous):
            raise ValueError("array must be contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(type(a))


def pinned_empty(shape, dtype=np.float64, order="F"):
    """numpy array backed by page-locked memory from gritas_b200_host_alloc (freed with the array)."""
    L = load()
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    nbytes = max(1, n * np.dtype(dtype).itemsize)
    p = L.gritas_b200_host_alloc(nbytes)
    if not p:
        raise MemoryError("gritas_b200_host_alloc failed")
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order=order)
    weakref.finalize(buf, L.gritas_b200_host_free, p)  # arr.base keeps buf alive
    return arr
