"""transpose_inspect / transpose — host-side mirror of the reference's CSR -> CSR transpose
(algorithms/transpose.hpp:8-13, algorithms/transpose_impl.hpp:9-60), calling the sm_100a
kernels through the C ABI (spblas_b200_transpose_inspect / spblas_b200_transpose).

Same argument meaning and error behaviour as the reference: `a` is m x n, `b` must be n x m
with values / colind arrays of at least a.nnz elements; afterwards b describes A^T with
b.nnz == a.nnz, rowptr zero-based, and inside every row of B the entries keep A's storage
order (the reference's scatter order) — structure and values are bit-identical to the
reference's result.  The inspect phase sorts the structure once; transpose(info, a, b) can
then be repeated after the values of `a` changed at the cost of one gather pass.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from .multiply import _stream_ptr, operation_info_t
from .views import csr_view, get_ultimate_base, index_type, value_type, _check_1d_cuda


def _check(a, b):
    a, b = get_ultimate_base(a), get_ultimate_base(b)
    if not isinstance(a, csr_view) or not isinstance(b, csr_view):
        raise TypeError("transpose: a and b must be csr_view")
    if a.shape[0] != b.shape[1] or a.shape[1] != b.shape[0]:
        raise ValueError("transpose: matrix dimensions are incompatible.")    # transpose_impl.hpp:17-21
    if b.values.numel() < a.nnz or b.colind.numel() < a.nnz:
        raise RuntimeError("transpose: Transpose ran out of memory.")        # transpose_impl.hpp:22-25
    if b.rowptr.numel() < b.shape[0] + 1:
        raise RuntimeError("transpose: Transpose ran out of memory.")
    for t, what in ((a.values, "a.values"), (a.rowptr, "a.rowptr"), (a.colind, "a.colind"),
                    (b.values, "b.values"), (b.rowptr, "b.rowptr"), (b.colind, "b.colind")):
        _check_1d_cuda(t, what)
    if (a.values.dtype, a.rowptr.dtype, a.colind.dtype) != (b.values.dtype, b.rowptr.dtype,
                                                            b.colind.dtype):
        raise RuntimeError("transpose: a and b must have the same scalar, index and offset types")
    return a, b


def _sig(a):
    return ("transpose", a.shape, a.nnz, a.rowptr.data_ptr(), a.colind.data_ptr(),
            a.rowptr.dtype, a.colind.dtype)


def _inspect(info: operation_info_t, a, b):
    a, b = _check(a, b)
    dev = a.values.device
    info._begin_inspect(_sig(a))
    plan = info._ensure(dev)
    with torch.cuda.device(dev):
        _cabi.lib().spblas_b200_plan_set_stream(plan, _stream_ptr(dev))
        st = _cabi.lib().spblas_b200_transpose_inspect(
            plan, a.shape[0], a.shape[1], a.nnz, a.rowptr.data_ptr(), a.colind.data_ptr(),
            index_type(a.rowptr), index_type(a.colind))
    _cabi.raise_for_status(st, info._err())
    info._sig = _sig(a)
    info.result_shape = b.shape
    info.result_nnz = a.nnz


def transpose_inspect(*args):
    """transpose_inspect(a, b) -> operation_info_t, or transpose_inspect(info, a, b) to
    re-inspect into an existing info (its device buffers are reused).  Analyses A's
    structure for B = A^T (a no-op in the CPU reference, transpose_impl.hpp:9-12; here:
    stable sort of the column indices carrying the storage position)."""
    if len(args) == 3 and isinstance(args[0], operation_info_t):
        _inspect(*args)
        return None
    if len(args) != 2:
        raise TypeError("transpose_inspect(a, b) or transpose_inspect(info, a, b)")
    info = operation_info_t()
    _inspect(info, *args)
    return info


def transpose(*args):
    """transpose(a, b) or transpose(info, a, b): B = A^T, both csr_view on the device."""
    if len(args) == 2:
        info, (a, b) = operation_info_t(), args
    elif len(args) == 3 and isinstance(args[0], operation_info_t):
        info, a, b = args
    else:
        raise TypeError("transpose(a, b) or transpose(info, a, b)")
    a_base, b_base = _check(a, b)
    if not info._select(_sig(a_base)):
        _inspect(info, a_base, b_base)
    dev = a_base.values.device
    with torch.cuda.device(dev):
        L = _cabi.lib()
        L.spblas_b200_plan_set_stream(info._plan, _stream_ptr(dev))
        st = L.spblas_b200_transpose(info._plan, value_type(a_base.values),
                                     a_base.values.data_ptr(), b_base.rowptr.data_ptr(),
                                     b_base.colind.data_ptr(), b_base.values.data_ptr())
    _cabi.raise_for_status(st, info._err())
    b_base.nnz = a_base.nnz          # b.update(..., a.size()), transpose_impl.hpp:52
    if len(args) == 2:
        info.close()
