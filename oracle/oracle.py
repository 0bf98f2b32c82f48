"""ctypes/numpy front end of oracle/liboracle.so (plain-C restatement of the reference
CPU path, oracle/spblas_oracle.c) and of oracle/_ref/libspblas_ref.so (the real
reference compiled from /root/reference by oracle/Makefile).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Parity status: pinned
(tests/test_oracle.py, tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libspblas_ref.so")

_VAL = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int32): "s32"}
_IDX = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}
_CT = {"f32": C.c_float, "f64": C.c_double, "s32": C.c_int32}


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible: gcc) and, where /root/reference exists,
    _ref/libspblas_ref.so."""
    if force or not os.path.exists(_ORACLE_SO) or (
        os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "spblas_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        _lib = C.CDLL(_ORACLE_SO)
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        if not have_ref():
            raise FileNotFoundError(
                f"{_REF_SO} missing: build it with `make -C oracle` where /root/reference exists"
            )
        _ref = C.CDLL(_REF_SO)
    return _ref


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _names(values, ind, ptr):
    return _VAL[values.dtype], _IDX[ind.dtype], _IDX[ptr.dtype]


def _scal(tn, v):
    return _CT[tn](v if v is not None else 0)


def _c(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


# ------------------------------------------------------------------ SpMV / SpMM
def spmv(fmt, shape, ptr, ind, values, x, alpha_a=None, alpha_x=None, impl="oracle",
         inspect=False):
    """y = A x following reference multiply_impl.hpp:33-53 (CSR rows / CSC columns).
    alpha_a / alpha_x model scaled(alpha, a) / scaled(alpha, x).  impl = "oracle" (C
    restatement) or "reference" (real spblas::multiply)."""
    m, n = shape
    ptr, ind, values, x = _c(ptr), _c(ind), _c(values), _c(x)
    tn, inn, on = _names(values, ind, ptr)
    assert x.dtype == values.dtype and x.shape == (n,)
    y = np.full(m, np.nan if values.dtype.kind == "f" else 7, dtype=values.dtype)
    if impl == "oracle":
        fn = getattr(lib(), f"oracle_{fmt}_spmv_{tn}_{inn}_{on}")
        fn.restype = None
        fn(C.c_int64(m), C.c_int64(n), _p(ptr), _p(ind), _p(values),
           C.c_int(alpha_a is not None), _scal(tn, alpha_a),
           C.c_int(alpha_x is not None), _scal(tn, alpha_x), _p(x), _p(y))
    else:
        fn = getattr(ref(), f"ref_{fmt}_spmv_{tn}_{inn}_{on}")
        fn.restype = None
        nnz = int(ptr[-1] - ptr[0]) if len(ptr) else 0
        fn(C.c_int64(m), C.c_int64(n), C.c_int64(nnz), _p(ptr), _p(ind), _p(values),
           C.c_int(alpha_a is not None), _scal(tn, alpha_a),
           C.c_int(alpha_x is not None), _scal(tn, alpha_x), _p(x), _p(y),
           C.c_int(int(inspect)))
    return y


def spmm(fmt, shape, ptr, ind, values, B, alpha_a=None, alpha_b=None, impl="oracle",
         inspect=False):
    """C = A B following reference multiply_impl.hpp:66-92, B/C row-major."""
    m, n = shape
    ptr, ind, values, B = _c(ptr), _c(ind), _c(values), _c(B)
    tn, inn, on = _names(values, ind, ptr)
    assert B.dtype == values.dtype and B.ndim == 2 and B.shape[0] == n
    k = B.shape[1]
    out = np.full((m, k), np.nan if values.dtype.kind == "f" else 7, dtype=values.dtype)
    if impl == "oracle":
        fn = getattr(lib(), f"oracle_{fmt}_spmm_{tn}_{inn}_{on}")
        fn.restype = None
        fn(C.c_int64(m), C.c_int64(n), C.c_int64(k), _p(ptr), _p(ind), _p(values),
           C.c_int(alpha_a is not None), _scal(tn, alpha_a),
           C.c_int(alpha_b is not None), _scal(tn, alpha_b),
           _p(B), C.c_int64(k), _p(out), C.c_int64(k))
    else:
        fn = getattr(ref(), f"ref_{fmt}_spmm_{tn}_{inn}_{on}")
        fn.restype = None
        nnz = int(ptr[-1] - ptr[0]) if len(ptr) else 0
        fn(C.c_int64(m), C.c_int64(n), C.c_int64(k), C.c_int64(nnz), _p(ptr), _p(ind),
           _p(values), C.c_int(alpha_a is not None), _scal(tn, alpha_a),
           C.c_int(alpha_b is not None), _scal(tn, alpha_b), _p(B), _p(out),
           C.c_int(int(inspect)))
    return out


def axpby(t, d, beta):
    """The 4-argument multiply's epilogue, y = t + beta * d with t = the 3-argument product
    (the convention of the reference's only 4-argument implementation,
    vendor/rocsparse/multiply_spgemm.hpp:69-118: beta = scaling factor of d): one multiply and
    one add per element, each rounded in the operands' type."""
    t, d = np.asarray(t), np.asarray(d)
    b = t.dtype.type(beta)
    return (t + b * d.astype(t.dtype)).astype(t.dtype)


def abs_rowsum(rowptr, colind, values, x, alpha=1.0):
    """s_i = sum_j |alpha a_ij x_j| in float64 (tolerance denominator)."""
    rowptr, colind, values, x = _c(rowptr), _c(colind), _c(values), _c(x)
    m = len(rowptr) - 1
    if values.dtype.kind != "f" or colind.dtype != np.int32:
        # generic numpy fallback for the rarely used type combinations
        prod = np.abs(alpha * values.astype(np.float64) * x.astype(np.float64)[colind])
        rows = np.repeat(np.arange(m), np.diff(rowptr.astype(np.int64)))
        return np.bincount(rows, weights=prod, minlength=m)
    tn, inn, on = _names(values, colind, rowptr)
    s = np.zeros(m, dtype=np.float64)
    fn = getattr(lib(), f"oracle_abs_rowsum_{tn}_{inn}_{on}")
    fn.restype = None
    fn(C.c_int64(m), _p(rowptr), _p(colind), _p(values), C.c_double(alpha), _p(x), _p(s))
    return s


# ------------------------------------------------------------------ inspect structures
HIST_BINS = 40


def rowlen_hist(rowptr, nbins=HIST_BINS):
    rp = _c(rowptr, np.int64)
    hist = np.zeros(nbins, dtype=np.int64)
    fn = lib().oracle_rowlen_hist
    fn.restype = C.c_int64
    mx = fn(C.c_int64(len(rp) - 1), _p(rp), C.c_int(nbins), _p(hist))
    return hist, int(mx)


def merge_partition(rowptr, tile_items):
    rp = _c(rowptr, np.int64)
    rows = len(rp) - 1
    nnz = int(rp[-1] - rp[0]) if rows > 0 else 0
    num_tiles = (rows + nnz + tile_items - 1) // tile_items
    starts = np.zeros(2 * (num_tiles + 1), dtype=np.int64)
    fn = lib().oracle_merge_partition
    fn.restype = None
    fn(C.c_int64(rows), _p(rp), C.c_int64(tile_items), C.c_int64(num_tiles), _p(starts))
    return starts.reshape(-1, 2)


def tile_uniform(rowptr, starts, max_len=8):
    rp = _c(rowptr, np.int64)
    st = _c(np.asarray(starts).reshape(-1), np.int64)
    nt = len(st) // 2 - 1
    out = np.zeros(max(nt, 1), dtype=np.int32)
    fn = lib().oracle_tile_uniform
    fn.restype = None
    fn(_p(rp), _p(st), C.c_int64(nt), C.c_int(max_len), _p(out))
    return out[:nt]


def row_segments(rowptr, seg):
    rp = _c(rowptr, np.int64)
    fn = lib().oracle_row_segments
    fn.restype = C.c_int64
    n = fn(C.c_int64(len(rp) - 1), _p(rp), C.c_int64(seg), None)
    out = np.zeros(3 * max(n, 1), dtype=np.int64)
    fn(C.c_int64(len(rp) - 1), _p(rp), C.c_int64(seg), _p(out))
    return out[: 3 * n].reshape(-1, 3)


def hub_columns(colind, n_cols, max_cols, min_count, by_popularity=False):
    """Definition of the hub-column analysis (csrc/hub.cu; no reference counterpart — the
    reference gathers x[j] per stored entry, multiply_impl.hpp:48-52): count the references
    per column, keep the columns referenced >= min_count times, order them by (count
    descending, column ascending), take the first max_cols, renumber them in ascending
    column order; a reference to hub number s is re-encoded as ~s.
    Returns (hub columns ascending, references to them, encoded colind).  by_popularity: the
    table of the global-memory kernel instead — the hubs keep the (count descending, column
    ascending) order, hub number s is the s-th most referenced column."""
    ci = np.asarray(colind).astype(np.int64)
    counts = np.bincount(ci, minlength=n_cols) if len(ci) else np.zeros(n_cols, np.int64)
    cand = np.nonzero(counts >= max(int(min_count), 1))[0]
    order = np.lexsort((cand, -counts[cand]))           # primary: count desc, then column asc
    top = cand[order][: int(max_cols)]
    refs = int(counts[top].sum())
    hubs = top if by_popularity else np.sort(top)
    slot = np.full(n_cols, -1, dtype=np.int64)
    slot[hubs] = np.arange(len(hubs))
    enc = np.where(slot[ci] >= 0, ~slot[ci], ci).astype(np.int32) if len(ci) else ci.astype(np.int32)
    return hubs.astype(np.int32), refs, enc


def csc_row_major_image(shape, colptr, rowind):
    m, n = shape
    cp, ri = _c(colptr, np.int64), _c(rowind, np.int64)
    nnz = len(ri)
    t_rowptr = np.zeros(m + 1, dtype=np.int64)
    t_colind = np.zeros(max(nnz, 1), dtype=np.int64)
    perm = np.zeros(max(nnz, 1), dtype=np.int64)
    fn = lib().oracle_csc_row_major_image
    fn.restype = None
    fn(C.c_int64(m), C.c_int64(n), _p(cp), _p(ri), _p(t_rowptr), _p(t_colind), _p(perm))
    return t_rowptr, t_colind[:nnz], perm[:nnz]


def trsv(m, rowptr, colind, values, b, upper=False, unit=False, alpha_a=None, alpha_b=None,
         impl="oracle", x0=None):
    """x = inv(tri(A)) b following reference algorithms/triangular_solve_impl.hpp:44-94
    (A: general square CSR; only the chosen triangle and, with an explicit diagonal, the
    diagonal entries are used).  alpha_a / alpha_b model scaled(alpha, a) / scaled(alpha, b).
    x0: initial contents of x (the reference reads x only at already solved positions)."""
    rowptr, colind, values, b = _c(rowptr), _c(colind), _c(values), _c(b)
    tn, inn, on = _names(values, colind, rowptr)
    assert b.dtype == values.dtype and b.shape == (m,)
    x = np.full(m, np.nan, dtype=values.dtype) if x0 is None else _c(x0).copy()
    args = (_p(rowptr), _p(colind), _p(values), C.c_int(int(upper)), C.c_int(int(unit)),
            C.c_int(alpha_a is not None), _scal(tn, alpha_a), C.c_int(alpha_b is not None),
            _scal(tn, alpha_b), _p(b), _p(x))
    if impl == "oracle":
        fn = getattr(lib(), f"oracle_csr_trsv_{tn}_{inn}_{on}")
        fn.restype = None
        fn(C.c_int64(m), *args)
    else:
        fn = getattr(ref(), f"ref_csr_trsv_{tn}_{inn}_{on}")
        fn.restype = None
        nnz = int(rowptr[-1] - rowptr[0]) if len(rowptr) else 0
        fn(C.c_int64(m), C.c_int64(nnz), *args)
    return x


def trsv_levels(m, rowptr, colind, upper=False):
    """Level of every row in the dependency graph of the chosen triangle: 0 for a row
    whose solve needs no other unknown, else 1 + the largest level among the unknowns it
    reads (definition used by csrc/trsv.cu; plain restatement, small cases only)."""
    rowptr, colind = _c(rowptr, np.int64), _c(colind, np.int64)
    level = np.zeros(m, dtype=np.int64)
    rows = range(m - 1, -1, -1) if upper else range(m)
    for i in rows:
        ks = colind[rowptr[i]:rowptr[i + 1]]
        deps = ks[ks > i] if upper else ks[ks < i]
        if len(deps):
            level[i] = level[deps].max() + 1
    return level


def transpose(shape, rowptr, colind, values, impl="oracle"):
    """B = A^T as CSR following reference algorithms/transpose_impl.hpp:14-53.
    Returns (b_values, b_rowptr, b_colind) with the dtypes of the inputs."""
    m, n = shape
    rowptr, colind, values = _c(rowptr), _c(colind), _c(values)
    tn, inn, on = _names(values, colind, rowptr)
    nnz = int(rowptr[-1] - rowptr[0]) if len(rowptr) else 0
    b_rowptr = np.full(n + 1, -1, dtype=rowptr.dtype)
    b_colind = np.full(max(nnz, 1), -1, dtype=colind.dtype)
    b_values = np.zeros(max(nnz, 1), dtype=values.dtype)
    if impl == "oracle":
        fn = getattr(lib(), f"oracle_csr_transpose_{tn}_{inn}_{on}")
        fn.restype = None
        fn(C.c_int64(m), C.c_int64(n), _p(rowptr), _p(colind), _p(values), _p(b_rowptr),
           _p(b_colind), _p(b_values))
    else:
        fn = getattr(ref(), f"ref_csr_transpose_{tn}_{inn}_{on}")
        fn.restype = None
        fn(C.c_int64(m), C.c_int64(n), C.c_int64(nnz), _p(rowptr), _p(colind), _p(values),
           _p(b_rowptr), _p(b_colind), _p(b_values))
    return b_values[:nnz], b_rowptr, b_colind[:nnz]


# ------------------------------------------------------------------ reference fixtures
def ref_generate_csr(m, n, nnz, seed=0, dtype=np.float32):
    """spblas::generate_csr<T, int32, int32> (reference backend/generate.hpp:106-120)."""
    tn = _VAL[np.dtype(dtype)]
    values = np.zeros(nnz, dtype=dtype)
    rowptr = np.zeros(m + 1, dtype=np.int32)
    colind = np.zeros(nnz, dtype=np.int32)
    fn = getattr(ref(), f"ref_generate_csr_{tn}_i32_i32")
    fn.restype = None
    fn(C.c_int64(m), C.c_int64(n), C.c_int64(nnz), C.c_int64(seed), _p(values), _p(rowptr),
       _p(colind))
    return values, rowptr, colind


def ref_generate_csc(m, n, nnz, seed=0, dtype=np.float32):
    """spblas::generate_csc<T, int32, int32> (reference backend/generate.hpp:131-138)."""
    tn = _VAL[np.dtype(dtype)]
    values = np.zeros(nnz, dtype=dtype)
    colptr = np.zeros(n + 1, dtype=np.int32)
    rowind = np.zeros(nnz, dtype=np.int32)
    fn = getattr(ref(), f"ref_generate_csc_{tn}_i32_i32")
    fn.restype = None
    fn(C.c_int64(m), C.c_int64(n), C.c_int64(nnz), C.c_int64(seed), _p(values), _p(colptr),
       _p(rowind))
    return values, colptr, rowind


def ref_generate_dense(m, n, seed=0, dtype=np.float32):
    """spblas::generate_dense<T> (reference backend/generate.hpp:170-182)."""
    tn = _VAL[np.dtype(dtype)]
    out = np.zeros((m, n), dtype=dtype)
    fn = getattr(ref(), f"ref_generate_dense_{tn}")
    fn.restype = None
    fn(C.c_int64(m), C.c_int64(n), C.c_int64(seed), _p(out))
    return out


def expect_eq_tolerance(t, u):
    """The reference tests' EXPECT_EQ_ (test/gtest/util.hpp:7-23): floating point within
    max(min_normal, 64 eps (|t| + |u|)); integers exactly.  Returns a bool array."""
    t, u = np.asarray(t), np.asarray(u)
    if t.dtype.kind != "f":
        return t == u
    fi = np.finfo(t.dtype)
    norm = np.minimum(np.abs(t).astype(np.float64) + np.abs(u).astype(np.float64), fi.max)
    abs_error = np.maximum(fi.tiny, 64 * fi.eps * norm)
    return np.abs(t.astype(np.float64) - u.astype(np.float64)) <= abs_error
