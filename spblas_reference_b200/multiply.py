"""multiply / multiply_inspect / multiply_execute — host-side mirror of the reference's
operator interface for the sparse-times-dense path, calling the sm_100a kernels through
the C ABI (include/spblas_b200.h).

Reference interface mirrored (names, argument meaning, error behaviour):
  multiply(a, x, y), multiply(info, a, x, y)             algorithms/multiply.hpp:15-26
  multiply_inspect(a, x, y), multiply_inspect(info, ..)  algorithms/multiply.hpp:9-13,29-33
  multiply_execute(info, a, x, y)                        README.md:33-47 (documented spelling)
  operation_info_t                                       detail/operation_info_t.hpp:28-104
Argument decoding follows vendor/cusparse/spmv_impl.hpp:29-40: peel views, reject
conjugated views, alpha = product of scaling factors, beta = 0.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _cabi
from .views import (csc_view, csr_view, get_scaling_factor, get_ultimate_base, has_matrix_opt,
                    index_type, is_conjugated, value_type, _check_1d_cuda)

_NP = {_cabi.F32: np.float32, _cabi.F64: np.float64, _cabi.S32: np.int32}


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _destroy_plan(plan):
    _cabi.lib().spblas_b200_plan_destroy(plan)


class operation_info_t:
    """Result of multiply_inspect: owns the backend plan (RAII, move-only in C++:
    include/spblas/vendor/b200/operation_state_t.hpp).  result_shape / result_nnz mirror
    detail/operation_info_t.hpp:30-36."""

    # One info may be inspected for more than one operand: the reference's notes inspect `a`
    # and `transposed(a)` with ONE operation_info_t and alternate the executes
    # (notes/spmv.hpp:12-22).  The current structure's plan is `_plan` / `_sig`; up to
    # _MAX_PARKED other structures keep their own plan, so that alternating executes switch
    # plans instead of re-inspecting (a CSC operand's inspect sorts the whole image).
    _MAX_PARKED = 3

    def __init__(self):
        self._plan = C.c_void_p()
        self._device = None
        self._sig = None
        self._parked = []            # [(sig, plan handle, device)], most recently used last
        self.result_shape = (0, 0)
        self.result_nnz = 0

    # -- plan lifetime ---------------------------------------------------------------
    def _ensure(self, device) -> C.c_void_p:
        if not self._plan:
            with torch.cuda.device(device):
                st = _cabi.lib().spblas_b200_plan_create(C.byref(self._plan), _stream_ptr(device))
            _cabi.raise_for_status(st, "plan_create")
            self._device = device
        return self._plan

    def _select(self, sig) -> bool:
        """Make the plan inspected for `sig` the current one; False if there is none."""
        if self._plan and self._sig == sig:
            return True
        for i, (s, plan, device) in enumerate(self._parked):
            if s == sig:
                del self._parked[i]
                self._park()
                self._plan, self._sig, self._device = plan, s, device
                return True
        return False

    def _park(self):
        """Set the current plan aside (its structure stays inspected) so that the next
        _ensure() creates a fresh one; the least recently used parked plan makes room."""
        if not self._plan:
            return
        if self._sig is None:        # never inspected: nothing worth keeping, reuse it
            return
        self._parked.append((self._sig, self._plan, self._device))
        self._plan, self._sig = C.c_void_p(), None
        while len(self._parked) > self._MAX_PARKED:
            _, old, _ = self._parked.pop(0)
            _destroy_plan(old)

    def _begin_inspect(self, sig):
        """Called by every inspect: reuse the plan of the same structure (a re-inspect),
        else park the current one."""
        if not self._select(sig):
            self._park()

    def close(self):
        if self._plan:
            _destroy_plan(self._plan)
            self._plan = C.c_void_p()
            self._sig = None
        for _, plan, _ in self._parked:
            _destroy_plan(plan)
        self._parked = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self) -> str:
        return _cabi.lib().spblas_b200_last_error(self._plan).decode()

    # -- metadata (SPBLAS_B200_Q_*) ---------------------------------------------------
    def _query_scalar(self, what: int) -> int:
        out = C.c_int64(0)
        st = _cabi.lib().spblas_b200_plan_query(self._plan, what, C.byref(out), 8, None)
        _cabi.raise_for_status(st, self._err())
        return int(out.value)

    def _query_array(self, what: int, dtype) -> np.ndarray:
        need = C.c_size_t(0)
        st = _cabi.lib().spblas_b200_plan_query(self._plan, what, None, 0, C.byref(need))
        _cabi.raise_for_status(st, self._err())
        buf = np.zeros(need.value // np.dtype(dtype).itemsize, dtype=dtype)
        if need.value:
            st = _cabi.lib().spblas_b200_plan_query(self._plan, what,
                                                    buf.ctypes.data_as(C.c_void_p),
                                                    need.value, None)
            _cabi.raise_for_status(st, self._err())
        return buf

    @property
    def num_tiles(self): return self._query_scalar(_cabi.Q_NUM_TILES)
    @property
    def tile_items(self): return self._query_scalar(_cabi.Q_TILE_ITEMS)
    @property
    def tile_starts(self): return self._query_array(_cabi.Q_TILE_STARTS, np.int64).reshape(-1, 2)
    @property
    def tile_uniform(self): return self._query_array(_cabi.Q_TILE_UNIFORM, np.int32)
    @property
    def rowlen_hist(self): return self._query_array(_cabi.Q_ROWLEN_HIST, np.int64)
    @property
    def max_row_len(self): return self._query_scalar(_cabi.Q_MAX_ROW_LEN)
    @property
    def empty_rows(self): return self._query_scalar(_cabi.Q_EMPTY_ROWS)
    @property
    def last_launches(self): return self._query_scalar(_cabi.Q_LAST_LAUNCHES)
    @property
    def total_launches(self): return self._query_scalar(_cabi.Q_TOTAL_LAUNCHES)
    @property
    def num_segments(self): return self._query_scalar(_cabi.Q_NUM_SEGMENTS)
    @property
    def segments(self): return self._query_array(_cabi.Q_SEGMENTS, np.int64).reshape(-1, 3)
    @property
    def spmv_variant(self): return self._query_scalar(_cabi.Q_SPMV_VARIANT)
    @property
    def spmm_variant(self): return self._query_scalar(_cabi.Q_SPMM_VARIANT)

    @property
    def trsv_levels(self): return self._query_scalar(_cabi.Q_TRSV_LEVELS)
    @property
    def trsv_sweeps(self): return self._query_scalar(_cabi.Q_TRSV_SWEEPS)
    @property
    def barrier_epoch(self): return self._query_scalar(_cabi.Q_BARRIER_EPOCH)
    @property
    def barrier_timeout(self): return self._query_scalar(_cabi.Q_BARRIER_TIMEOUT)

    # -- hub columns (include/spblas_b200.h: spblas_b200_plan_set_hub) ---------------------
    @property
    def hub_count(self): return self._query_scalar(_cabi.Q_HUB_COUNT)
    @property
    def hub_refs(self): return self._query_scalar(_cabi.Q_HUB_REFS)
    @property
    def hub_cols(self): return self._query_array(_cabi.Q_HUB_COLS, np.int32)
    @property
    def hub_colind(self): return self._query_array(_cabi.Q_HUB_COLIND, np.int32)

    def set_hub(self, enable: bool = True, max_cols: int = 0, min_count: int = 0):
        """Let the general SpMV path keep x at the most referenced columns in shared memory
        (power-law matrices).  max_cols / min_count: 0 = the backend's defaults, -1 = keep."""
        st = _cabi.lib().spblas_b200_plan_set_hub(self._plan, 1 if enable else 0,
                                                  int(max_cols), int(min_count))
        _cabi.raise_for_status(st, self._err())

    def force_spmv_variant(self, variant: int):
        """Tuning / test knob: 0 merge-tile, 1 pipelined, 2 warp-stream, 3 hub table in shared
        memory, 4 hub table in global memory, -1 automatic (spblas_b200_plan_force_variant)."""
        st = _cabi.lib().spblas_b200_plan_force_variant(self._plan, int(variant))
        _cabi.raise_for_status(st, self._err())

    # -- fused exchange (include/spblas_b200.h: set_scatter / set_barrier) ---------------
    def set_scatter(self, dsts=(), multicast: bool = False):
        """dsts: (device address of this block's row 0 in the destination, row_begin, row_end)
        triples; every following SpMV execute on this plan also stores those rows there."""
        n = len(dsts)
        ptrs = (C.c_void_p * max(n, 1))(*[int(d[0]) for d in dsts])
        lo = (C.c_int64 * max(n, 1))(*[int(d[1]) for d in dsts])
        hi = (C.c_int64 * max(n, 1))(*[int(d[2]) for d in dsts])
        st = _cabi.lib().spblas_b200_plan_set_scatter(self._plan, n, ptrs, lo, hi,
                                                      1 if multicast else 0)
        _cabi.raise_for_status(st, self._err())

    def set_barrier(self, remote_slots=(), local_slots=()):
        """remote_slots[q]: address of this rank's flag word on peer q; local_slots[q]: address
        of the word peer q writes here.  Empty: no barrier."""
        n = len(remote_slots)
        assert n == len(local_slots)
        rs = (C.c_void_p * max(n, 1))(*[int(a) for a in remote_slots])
        ls = (C.c_void_p * max(n, 1))(*[int(a) for a in local_slots])
        st = _cabi.lib().spblas_b200_plan_set_barrier(self._plan, n, rs, ls)
        _cabi.raise_for_status(st, self._err())

    def effective_csr(self, off_dtype, idx_dtype):
        """(rowptr, colind, perm) of the row-major structure the kernels run on."""
        return (self._query_array(_cabi.Q_CSR_ROWPTR, off_dtype),
                self._query_array(_cabi.Q_CSR_COLIND, idx_dtype),
                self._query_array(_cabi.Q_CSR_PERM, off_dtype))


# ---------------------------------------------------------------------------------------
def _decode_matrix(a):
    if is_conjugated(a):
        raise RuntimeError("b200 backend does not support conjugated views.")
    base = get_ultimate_base(a)
    if isinstance(base, csr_view):
        fmt, ptr, ind = _cabi.CSR, base.rowptr, base.colind
    elif isinstance(base, csc_view):
        fmt, ptr, ind = _cabi.CSC, base.colptr, base.rowind
    else:
        raise TypeError("multiply: A must be a csr_view or csc_view (possibly wrapped)")
    _check_1d_cuda(base.values, "A.values")
    _check_1d_cuda(ptr, "A offsets")
    _check_1d_cuda(ind, "A indices")
    return base, fmt, ptr, ind


def _is_matrix(t) -> bool:
    return isinstance(t, torch.Tensor) and t.dim() == 2


def _check_shapes(a_base, x_base, y):
    m, n = a_base.shape
    if _is_matrix(y) != _is_matrix(x_base):
        raise TypeError("multiply: x and y must both be vectors or both be matrices")
    if _is_matrix(y):
        # reference multiply_impl.hpp:70-75
        if m != y.shape[0] or x_base.shape[1] != y.shape[1] or n != x_base.shape[0]:
            raise ValueError("multiply: matrix dimensions are incompatible.")
    else:
        # reference multiply_impl.hpp:37-41
        if m != y.shape[0] or n != x_base.shape[0]:
            raise ValueError("multiply: matrix and vector dimensions are incompatible.")


def _row_major(t: torch.Tensor, what: str):
    """mdspan_row_major contract (detail/mdspan.hpp:38-41): unit column stride; the row
    stride is the leading dimension."""
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise RuntimeError(f"multiply: {what} must be row-major (layout_right)")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return int(max(ld, t.shape[1]))


def _signature(fmt, a_base, ptr, ind):
    return (fmt, a_base.shape, a_base.nnz, ptr.data_ptr(), ind.data_ptr(), ptr.dtype, ind.dtype)


def _inspect(info: operation_info_t, a, x, y, flags=_cabi.INSPECT_DEFAULT):
    a_base, fmt, ptr, ind = _decode_matrix(a)
    x_base = get_ultimate_base(x)
    _check_shapes(a_base, x_base, y)
    dev = a_base.values.device
    info._begin_inspect(_signature(fmt, a_base, ptr, ind))
    plan = info._ensure(dev)
    k_hint = int(y.shape[1]) if _is_matrix(y) else 1
    with torch.cuda.device(dev):
        _cabi.lib().spblas_b200_plan_set_stream(plan, _stream_ptr(dev))
        st = _cabi.lib().spblas_b200_inspect(
            plan, fmt, a_base.shape[0], a_base.shape[1], a_base.nnz, ptr.data_ptr(),
            ind.data_ptr(), index_type(ptr), index_type(ind), k_hint, flags)
    _cabi.raise_for_status(st, info._err())
    if has_matrix_opt(a) and flags == _cabi.INSPECT_DEFAULT:
        # ... and structure-only state: x at the most referenced columns in shared memory,
        # decided on the first product (spblas_b200_plan_set_hub; bit-identical results)
        st = _cabi.lib().spblas_b200_plan_set_hub(plan, 1, -1, -1)
        _cabi.raise_for_status(st, info._err())
    if has_matrix_opt(a) and fmt == _cabi.CSC and flags == _cabi.INSPECT_DEFAULT:
        # matrix_opt: the backend may keep optimised, value-dependent state (the reference's
        # oneMKL backend calls optimize_gemv under the same condition) — here the values
        # gathered once into the order of the row-major image
        with torch.cuda.device(dev):
            st = _cabi.lib().spblas_b200_plan_cache_values(plan, value_type(a_base.values),
                                                           a_base.values.data_ptr())
        _cabi.raise_for_status(st, info._err())
    info._sig = _signature(fmt, a_base, ptr, ind)
    info.result_shape = tuple(y.shape) if _is_matrix(y) else (int(y.shape[0]), 1)
    info.result_nnz = int(y.numel())


def multiply_inspect(*args):
    """multiply_inspect(a, x, y) -> operation_info_t, or multiply_inspect(info, a, x, y).
    Runs the GPU inspect phase: row-length histogram, merge-path partition, CSC image,
    SpMM row segments (csrc/inspect.cu)."""
    if len(args) == 3:
        info = operation_info_t()
        _inspect(info, *args)
        return info
    if len(args) == 4 and isinstance(args[0], operation_info_t):
        _inspect(args[0], *args[1:])
        return None
    raise TypeError("multiply_inspect(a, x, y) or multiply_inspect(info, a, x, y)")


def _decode_addend(d, y, vt):
    """The 4-argument form's d (vendor/rocsparse/multiply_spgemm.hpp:69-118 convention):
    beta = scaling factor of d (1 if none); d has y's shape and type; it may BE y."""
    if is_conjugated(d):
        raise RuntimeError("b200 backend does not support conjugated views.")
    d_base = get_ultimate_base(d)
    if not isinstance(d_base, torch.Tensor) or not d_base.is_cuda:
        raise RuntimeError("d must live in device memory (no CPU path)")
    if tuple(d_base.shape) != tuple(y.shape):
        raise ValueError("multiply: matrix and vector dimensions are incompatible."
                         if not _is_matrix(y) else "multiply: matrix dimensions are incompatible.")
    if d_base.dtype != y.dtype:
        raise RuntimeError("b200 backend needs A, x, y and d of one scalar type")
    beta = get_scaling_factor(d)
    beta_np = np.array([1 if beta is None else beta], dtype=_NP[vt])
    return d_base, beta_np


def _execute(info: Optional[operation_info_t], a, x, y, d=None):
    a_base, fmt, ptr, ind = _decode_matrix(a)
    if is_conjugated(x) or is_conjugated(y):
        raise RuntimeError("b200 backend does not support conjugated views.")
    x_base = get_ultimate_base(x)
    if not isinstance(y, torch.Tensor):
        raise TypeError("multiply: the output must be a plain tensor (no views)")
    _check_shapes(a_base, x_base, y)
    for t, what in ((x_base, "x"), (y, "y")):
        if not t.is_cuda:
            raise RuntimeError(f"{what} must live in device memory (no CPU path)")
    vt = value_type(a_base.values)
    if x_base.dtype != a_base.values.dtype or y.dtype != a_base.values.dtype:
        raise RuntimeError("b200 backend needs A, x and y of one scalar type")
    alpha = get_scaling_factor(a, x)
    alpha_np = np.array([1 if alpha is None else alpha], dtype=_NP[vt])
    alpha_p = alpha_np.ctypes.data_as(C.c_void_p)
    dev = a_base.values.device
    L = _cabi.lib()
    m, n = a_base.shape
    d_base = beta_np = beta_p = None
    if d is not None:
        d_base, beta_np = _decode_addend(d, y, vt)
        beta_p = beta_np.ctypes.data_as(C.c_void_p)

    with torch.cuda.device(dev):
        stream = _stream_ptr(dev)
        if info is None:
            # no operation_info_t: one-shot entry points (thread-local plan; a structure seen
            # before is reused after a device-side check of its offsets array)
            if _is_matrix(y):
                common = (stream, fmt, m, n, a_base.nnz, ptr.data_ptr(), ind.data_ptr(),
                          index_type(ptr), index_type(ind), vt, alpha_p, a_base.values.data_ptr(),
                          x_base.data_ptr(), _row_major(x_base, "B"))
                if d is None:
                    st = L.spblas_b200_spmm_once(*common, y.data_ptr(), _row_major(y, "C"),
                                                 int(y.shape[1]))
                else:
                    st = L.spblas_b200_spmm_axpby_once(*common, beta_p, d_base.data_ptr(),
                                                       _row_major(d_base, "D"), y.data_ptr(),
                                                       _row_major(y, "C"), int(y.shape[1]))
            else:
                _check_1d_cuda(x_base, "x")
                _check_1d_cuda(y, "y")
                common = (stream, fmt, m, n, a_base.nnz, ptr.data_ptr(), ind.data_ptr(),
                          index_type(ptr), index_type(ind), vt, alpha_p, a_base.values.data_ptr(),
                          x_base.data_ptr())
                if d is None:
                    st = L.spblas_b200_spmv_once(*common, y.data_ptr())
                else:
                    _check_1d_cuda(d_base, "d")
                    st = L.spblas_b200_spmv_axpby_once(*common, beta_p, d_base.data_ptr(),
                                                       y.data_ptr())
            _cabi.raise_for_status(st, L.spblas_b200_last_error_once().decode())
            return

        if not info._select(_signature(fmt, a_base, ptr, ind)):
            # first use of this info for this matrix: inspect lazily, like
            # vendor/cusparse/spmv_impl.hpp:43-55 creates its state on first use
            _inspect(info, a, x, y)
        plan = info._plan
        L.spblas_b200_plan_set_stream(plan, stream)
        if _is_matrix(y):
            if d is None:
                st = L.spblas_b200_spmm(plan, vt, alpha_p, a_base.values.data_ptr(),
                                        x_base.data_ptr(), _row_major(x_base, "B"), y.data_ptr(),
                                        _row_major(y, "C"), int(y.shape[1]))
            else:
                st = L.spblas_b200_spmm_axpby(plan, vt, alpha_p, a_base.values.data_ptr(),
                                              x_base.data_ptr(), _row_major(x_base, "B"), beta_p,
                                              d_base.data_ptr(), _row_major(d_base, "D"),
                                              y.data_ptr(), _row_major(y, "C"), int(y.shape[1]))
        else:
            _check_1d_cuda(x_base, "x")
            _check_1d_cuda(y, "y")
            if d is None:
                st = L.spblas_b200_spmv(plan, vt, alpha_p, a_base.values.data_ptr(),
                                        x_base.data_ptr(), y.data_ptr())
            else:
                _check_1d_cuda(d_base, "d")
                st = L.spblas_b200_spmv_axpby(plan, vt, alpha_p, a_base.values.data_ptr(),
                                              x_base.data_ptr(), beta_p, d_base.data_ptr(),
                                              y.data_ptr())
        _cabi.raise_for_status(st, info._err())


def multiply(*args):
    """multiply(a, x, y) or multiply(info, a, x, y): y = alpha * A * x (SpMV) or
    C = alpha * A * B (SpMM, row-major 2-D tensors); y / C is overwritten (beta = 0).
    multiply(a, x, y, d) or multiply(info, a, x, y, d): the 4-argument form sketched in the
    reference's notes/matrices.hpp and implemented there for rocSPARSE SpGEMM
    (vendor/rocsparse/multiply_spgemm.hpp:69-118): y = alpha * A * x + beta * d with
    beta = the scaling factor of d (multiply(a, x, y, scaled(beta, d)); 1 if d is not scaled);
    d may be y itself.  The addend is fused into the kernels' single store per row."""
    if args and isinstance(args[0], operation_info_t):
        if len(args) in (4, 5):
            return _execute(*args)
    elif len(args) in (3, 4):
        return _execute(None, *args)
    raise TypeError("multiply(a, x, y[, d]) or multiply(info, a, x, y[, d])")


def multiply_execute(info: operation_info_t, a, x, y, d=None):
    """The execute phase under the name the reference documents (README.md:46,
    notes/spmv.hpp:20-22); identical to multiply(info, a, x, y[, d])."""
    if not isinstance(info, operation_info_t):
        raise TypeError("multiply_execute(info, a, x, y[, d])")
    return _execute(info, a, x, y, d)


def multiply_execute_host(info: operation_info_t, a, x_host: torch.Tensor, y_host: torch.Tensor):
    """y_host = alpha * A * x_host for HOST vectors (pinned memory for asynchronous copies);
    A and the inspected plan stay on the device.  One C-ABI call (spblas_b200_spmv_host)
    that pipelines upload, kernels and download chunk by chunk over the plan's tiles
    (csrc/host_exec.cu); bit-identical to multiply_execute on device vectors.  Complete in
    the order of the current stream: synchronise before reading y_host.  The device staging
    vectors are owned by `info` and reused."""
    if not isinstance(info, operation_info_t):
        raise TypeError("multiply_execute_host(info, a, x_host, y_host)")
    a_base, fmt, ptr, ind = _decode_matrix(a)
    if is_conjugated(x_host) or is_conjugated(y_host):
        raise RuntimeError("b200 backend does not support conjugated views.")
    x_base = get_ultimate_base(x_host)
    if not isinstance(y_host, torch.Tensor) or _is_matrix(y_host):
        raise TypeError("multiply_execute_host: SpMV only, plain 1-D host tensors")
    _check_shapes(a_base, x_base, y_host)
    for t, what in ((x_base, "x_host"), (y_host, "y_host")):
        if t.is_cuda or not t.is_contiguous():
            raise RuntimeError(f"{what} must be a contiguous host tensor")
    vt = value_type(a_base.values)
    if x_base.dtype != a_base.values.dtype or y_host.dtype != a_base.values.dtype:
        raise RuntimeError("b200 backend needs A, x and y of one scalar type")
    alpha = get_scaling_factor(a, x_host)
    alpha_np = np.array([1 if alpha is None else alpha], dtype=_NP[vt])
    dev = a_base.values.device
    m, n = a_base.shape
    with torch.cuda.device(dev):
        if not info._select(_signature(fmt, a_base, ptr, ind)):
            _inspect(info, a, torch.empty(n, dtype=x_base.dtype, device=dev),
                     torch.empty(m, dtype=x_base.dtype, device=dev))
        stage = getattr(info, "_stage", None)
        if stage is None or stage[0].dtype != x_base.dtype or stage[0].numel() < n \
                or stage[1].numel() < m:
            stage = (torch.empty(max(n, 1), dtype=x_base.dtype, device=dev),
                     torch.empty(max(m, 1), dtype=x_base.dtype, device=dev))
            info._stage = stage
        L = _cabi.lib()
        L.spblas_b200_plan_set_stream(info._plan, _stream_ptr(dev))
        st = L.spblas_b200_spmv_host(info._plan, vt, alpha_np.ctypes.data_as(C.c_void_p),
                                     a_base.values.data_ptr(), x_base.data_ptr(),
                                     y_host.data_ptr(), stage[0].data_ptr(), stage[1].data_ptr())
    _cabi.raise_for_status(st, info._err())
