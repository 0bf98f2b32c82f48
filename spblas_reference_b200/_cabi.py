"""ctypes binding of libspblas_b200.so — the C ABI declared in include/spblas_b200.h.

This is the only door from Python into the product: hand-written sm_100a kernels
behind `extern "C"`.  There is no CPU fallback: if the library is missing or does
not load, importing the host API raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
# SPBLAS_B200_LIB: another build of the same library (A/B measurements of one experiment)
LIB_PATH = os.environ.get("SPBLAS_B200_LIB") or os.path.join(_PKG, "libspblas_b200.so")

# ---- enums (include/spblas_b200.h) ------------------------------------------------
SUCCESS, INVALID_ARGUMENT, SHAPE_MISMATCH, NOT_SUPPORTED = 0, 1, 2, 3
ALLOC_FAILED, CUDA_ERROR, NOT_INSPECTED, INVALID_STRUCTURE = 4, 5, 6, 7
CSR, CSC = 0, 1
I32, I64 = 0, 1
F32, F64, S32 = 0, 1, 2
INSPECT_DEFAULT, INSPECT_LIGHT = 0, 1
(Q_NUM_TILES, Q_TILE_ITEMS, Q_TILE_STARTS, Q_ROWLEN_HIST, Q_MAX_ROW_LEN, Q_EMPTY_ROWS,
 Q_SPMV_VARIANT, Q_LAST_LAUNCHES, Q_TOTAL_LAUNCHES, Q_CSR_ROWPTR, Q_CSR_COLIND,
 Q_CSR_PERM, Q_NUM_SEGMENTS, Q_SEGMENTS, Q_SPMM_VARIANT, Q_TILE_UNIFORM,
 Q_BARRIER_EPOCH, Q_BARRIER_TIMEOUT, Q_TRSV_LEVELS, Q_TRSV_SWEEPS,
 Q_HUB_COUNT, Q_HUB_REFS, Q_HUB_COLS, Q_HUB_COLIND) = range(24)
MAX_PEERS = 8
HIST_BINS = 40

# every symbol include/spblas_b200.h declares (tests/test_cabi_symbols.py checks the
# header and this list against the built library)
SYMBOLS = (
    "spblas_b200_plan_create", "spblas_b200_plan_destroy", "spblas_b200_plan_set_stream",
    "spblas_b200_inspect", "spblas_b200_spmv", "spblas_b200_spmm",
    "spblas_b200_spmv_once", "spblas_b200_spmm_once", "spblas_b200_plan_query",
    "spblas_b200_last_error", "spblas_b200_last_error_once", "spblas_b200_status_string",
    "spblas_b200_version", "spblas_b200_plan_force_variant",
    "spblas_b200_plan_set_scatter", "spblas_b200_plan_set_barrier",
    "spblas_b200_spmv_host", "spblas_b200_probe_gather",
    "spblas_b200_transpose_inspect", "spblas_b200_transpose",
    "spblas_b200_plan_cache_values", "spblas_b200_trsv_inspect", "spblas_b200_trsv",
    "spblas_b200_plan_set_hub",
    "spblas_b200_spmv_axpby", "spblas_b200_spmm_axpby",
    "spblas_b200_spmv_axpby_once", "spblas_b200_spmm_axpby_once", "spblas_b200_once_release",
)


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_PKG, "csrc"), "-j8"],
                       capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libspblas_b200.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the B200 backend has no CPU fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C spblas_reference_b200/csrc`.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.spblas_b200_plan_create.argtypes = [C.POINTER(vp), vp]
    L.spblas_b200_plan_create.restype = i32
    L.spblas_b200_plan_destroy.argtypes = [vp]
    L.spblas_b200_plan_destroy.restype = None
    L.spblas_b200_plan_set_stream.argtypes = [vp, vp]
    L.spblas_b200_plan_set_stream.restype = i32
    L.spblas_b200_plan_force_variant.argtypes = [vp, i32]
    L.spblas_b200_plan_force_variant.restype = i32
    L.spblas_b200_plan_set_hub.argtypes = [vp, i32, i64, i64]
    L.spblas_b200_plan_set_hub.restype = i32
    L.spblas_b200_plan_set_scatter.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64),
                                               C.POINTER(i64), i32]
    L.spblas_b200_plan_set_scatter.restype = i32
    L.spblas_b200_plan_set_barrier.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    L.spblas_b200_plan_set_barrier.restype = i32
    L.spblas_b200_inspect.argtypes = [vp, i32, i64, i64, i64, vp, vp, i32, i32, i64, i32]
    L.spblas_b200_inspect.restype = i32
    L.spblas_b200_spmv.argtypes = [vp, i32, vp, vp, vp, vp]
    L.spblas_b200_spmv.restype = i32
    L.spblas_b200_spmv_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.spblas_b200_spmv_host.restype = i32
    L.spblas_b200_trsv_inspect.argtypes = [vp, i64, i64, vp, vp, i32, i32, i32, i32]
    L.spblas_b200_trsv_inspect.restype = i32
    L.spblas_b200_trsv.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.spblas_b200_trsv.restype = i32
    L.spblas_b200_plan_cache_values.argtypes = [vp, i32, vp]
    L.spblas_b200_plan_cache_values.restype = i32
    L.spblas_b200_transpose_inspect.argtypes = [vp, i64, i64, i64, vp, vp, i32, i32]
    L.spblas_b200_transpose_inspect.restype = i32
    L.spblas_b200_transpose.argtypes = [vp, i32, vp, vp, vp, vp]
    L.spblas_b200_transpose.restype = i32
    L.spblas_b200_probe_gather.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, i32]
    L.spblas_b200_probe_gather.restype = i32
    L.spblas_b200_spmm.argtypes = [vp, i32, vp, vp, vp, i64, vp, i64, i64]
    L.spblas_b200_spmm.restype = i32
    L.spblas_b200_spmv_once.argtypes = [vp, i32, i64, i64, i64, vp, vp, i32, i32, i32, vp,
                                        vp, vp, vp]
    L.spblas_b200_spmv_once.restype = i32
    L.spblas_b200_spmm_once.argtypes = [vp, i32, i64, i64, i64, vp, vp, i32, i32, i32, vp,
                                        vp, vp, i64, vp, i64, i64]
    L.spblas_b200_spmm_once.restype = i32
    L.spblas_b200_spmv_axpby.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.spblas_b200_spmv_axpby.restype = i32
    L.spblas_b200_spmm_axpby.argtypes = [vp, i32, vp, vp, vp, i64, vp, vp, i64, vp, i64, i64]
    L.spblas_b200_spmm_axpby.restype = i32
    L.spblas_b200_spmv_axpby_once.argtypes = [vp, i32, i64, i64, i64, vp, vp, i32, i32, i32, vp,
                                              vp, vp, vp, vp, vp]
    L.spblas_b200_spmv_axpby_once.restype = i32
    L.spblas_b200_spmm_axpby_once.argtypes = [vp, i32, i64, i64, i64, vp, vp, i32, i32, i32, vp,
                                              vp, vp, i64, vp, vp, i64, vp, i64, i64]
    L.spblas_b200_spmm_axpby_once.restype = i32
    L.spblas_b200_once_release.argtypes = []
    L.spblas_b200_once_release.restype = None
    L.spblas_b200_plan_query.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.spblas_b200_plan_query.restype = i32
    L.spblas_b200_last_error.argtypes = [vp]
    L.spblas_b200_last_error.restype = C.c_char_p
    L.spblas_b200_last_error_once.argtypes = []
    L.spblas_b200_last_error_once.restype = C.c_char_p
    L.spblas_b200_status_string.argtypes = [i32]
    L.spblas_b200_status_string.restype = C.c_char_p
    L.spblas_b200_version.argtypes = []
    L.spblas_b200_version.restype = i32
    _lib = L
    return L


def raise_for_status(status: int, message: str) -> None:
    """Map a C-ABI status onto the exception the reference throws for the same
    condition (include/spblas/vendor/b200/exception.hpp does the same in C++):
    shape mismatch -> ValueError (std::invalid_argument, multiply_impl.hpp:37-41),
    allocation -> MemoryError (std::bad_alloc, vendor/cusparse/cuda_allocator.hpp:60-64),
    everything else -> RuntimeError (vendor/cusparse/exception.hpp:13-21)."""
    if status == SUCCESS:
        return
    what = lib().spblas_b200_status_string(status).decode()
    text = f"{what}: {message}" if message else what
    if status in (SHAPE_MISMATCH, INVALID_ARGUMENT):
        raise ValueError(text)
    if status == ALLOC_FAILED:
        raise MemoryError(text)
    raise RuntimeError(text)
