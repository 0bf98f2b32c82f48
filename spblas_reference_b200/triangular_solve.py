"""triangular_solve_inspect / triangular_solve — host-side mirror of the reference's SpTRSV
interface (algorithms/triangular_solve.hpp:8-19, algorithms/triangular_solve_impl.hpp:14-107,
detail/triangular_types.hpp), calling the sm_100a kernels through the C ABI
(spblas_b200_trsv_inspect / spblas_b200_trsv, csrc/trsv.cu).

    x = inv(tri(A)) b        A: square csr_view on the device (possibly scaled / matrix_opt)
                             uplo: lower_triangle | upper_triangle
                             diag: explicit_diagonal | implicit_unit_diagonal
                             b:    device vector (possibly scaled(alpha, b)),  x: device vector

Only the chosen triangle (and the stored diagonal, unless implicit) of A is used, as in the
reference.  The result is bit-identical to the reference's serial loop.  With
explicit_diagonal every row must store a diagonal entry (RuntimeError at inspect otherwise:
the reference divides by the previous row's diagonal there).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from .multiply import _NP, _stream_ptr, operation_info_t
from .views import (csr_view, get_ultimate_base, index_type, is_conjugated, scaled_view,
                    conjugated_view, matrix_opt, value_type, _check_1d_cuda)


class upper_triangle_t:
    pass


class lower_triangle_t:
    pass


class implicit_unit_diagonal_t:
    pass


class explicit_diagonal_t:
    pass


upper_triangle, lower_triangle = upper_triangle_t(), lower_triangle_t()
implicit_unit_diagonal, explicit_diagonal = implicit_unit_diagonal_t(), explicit_diagonal_t()


def _own_scaling(t):
    """Product of the scaled_view factors of ONE operand (None if it has none)."""
    out = None
    while isinstance(t, (scaled_view, conjugated_view, matrix_opt)):
        if isinstance(t, scaled_view):
            out = t.alpha if out is None else out * t.alpha
        t = t.base
    return out


def _decode(a, uplo, diag, b, x):
    if not isinstance(uplo, (upper_triangle_t, lower_triangle_t)):
        raise TypeError("triangular_solve: uplo must be upper_triangle or lower_triangle")
    if not isinstance(diag, (implicit_unit_diagonal_t, explicit_diagonal_t)):
        raise TypeError("triangular_solve: diag must be explicit_diagonal or implicit_unit_diagonal")
    if is_conjugated(a) or is_conjugated(b) or is_conjugated(x):
        raise RuntimeError("b200 backend does not support conjugated views.")
    a_base, b_base = get_ultimate_base(a), get_ultimate_base(b)
    if not isinstance(a_base, csr_view):
        raise TypeError("triangular_solve: A must be a csr_view (possibly wrapped)")
    if not isinstance(x, torch.Tensor):
        raise TypeError("triangular_solve: the output must be a plain tensor (no views)")
    m, n = a_base.shape
    if m != n:                                   # assert in triangular_solve_impl.hpp:21,50
        raise ValueError("triangular_solve: matrix must be square.")
    if x.shape != (n,) or b_base.shape != (m,):   # assert in triangular_solve_impl.hpp:52-53
        raise ValueError("triangular_solve: matrix and vector dimensions are incompatible.")
    for t, what in ((a_base.values, "A.values"), (a_base.rowptr, "A.rowptr"),
                    (a_base.colind, "A.colind"), (b_base, "b"), (x, "x")):
        _check_1d_cuda(t, what)
    if b_base.dtype != a_base.values.dtype or x.dtype != a_base.values.dtype:
        raise RuntimeError("b200 backend needs A, b and x of one scalar type")
    return a_base, b_base


def _sig(a_base, uplo, diag):
    return ("trsv", a_base.shape, a_base.nnz, a_base.rowptr.data_ptr(), a_base.colind.data_ptr(),
            a_base.rowptr.dtype, a_base.colind.dtype, type(uplo).__name__, type(diag).__name__)


def _inspect(info: operation_info_t, a, uplo, diag, b, x):
    a_base, _ = _decode(a, uplo, diag, b, x)
    dev = a_base.values.device
    info._begin_inspect(_sig(a_base, uplo, diag))
    plan = info._ensure(dev)
    with torch.cuda.device(dev):
        _cabi.lib().spblas_b200_plan_set_stream(plan, _stream_ptr(dev))
        st = _cabi.lib().spblas_b200_trsv_inspect(
            plan, a_base.shape[0], a_base.nnz, a_base.rowptr.data_ptr(),
            a_base.colind.data_ptr(), index_type(a_base.rowptr), index_type(a_base.colind),
            int(isinstance(uplo, upper_triangle_t)), int(isinstance(diag, implicit_unit_diagonal_t)))
    _cabi.raise_for_status(st, info._err())
    info._sig = _sig(a_base, uplo, diag)
    info.result_shape = (int(x.shape[0]), 1)
    info.result_nnz = int(x.numel())


def triangular_solve_inspect(*args):
    """triangular_solve_inspect(a, uplo, diag, b, x) -> operation_info_t, or
    triangular_solve_inspect(info, a, uplo, diag, b, x).  GPU level-set analysis of the
    chosen triangle (a no-op in the CPU reference, triangular_solve_impl.hpp:14-41)."""
    if len(args) == 5:
        info = operation_info_t()
        _inspect(info, *args)
        return info
    if len(args) == 6 and isinstance(args[0], operation_info_t):
        _inspect(*args)
        return None
    raise TypeError("triangular_solve_inspect(a, uplo, diag, b, x) or (info, a, uplo, diag, b, x)")


def triangular_solve(*args):
    """triangular_solve(a, uplo, diag, b, x) or triangular_solve(info, a, uplo, diag, b, x)."""
    if len(args) == 5:
        info, rest, own = operation_info_t(), args, True
    elif len(args) == 6 and isinstance(args[0], operation_info_t):
        info, rest, own = args[0], args[1:], False
    else:
        raise TypeError("triangular_solve(a, uplo, diag, b, x) or (info, a, uplo, diag, b, x)")
    a, uplo, diag, b, x = rest
    a_base, b_base = _decode(a, uplo, diag, b, x)
    if not info._select(_sig(a_base, uplo, diag)):
        _inspect(info, a, uplo, diag, b, x)
    vt = value_type(a_base.values)
    if vt == _cabi.S32:
        raise RuntimeError("b200 backend: triangular_solve needs floating-point scalars")
    aa, ab = _own_scaling(a), _own_scaling(b)
    aa_np = None if aa is None else np.array([aa], dtype=_NP[vt])
    ab_np = None if ab is None else np.array([ab], dtype=_NP[vt])
    dev = a_base.values.device
    with torch.cuda.device(dev):
        L = _cabi.lib()
        L.spblas_b200_plan_set_stream(info._plan, _stream_ptr(dev))
        st = L.spblas_b200_trsv(info._plan, vt,
                                None if aa_np is None else aa_np.ctypes.data_as(C.c_void_p),
                                None if ab_np is None else ab_np.ctypes.data_as(C.c_void_p),
                                a_base.values.data_ptr(), b_base.data_ptr(), x.data_ptr())
    _cabi.raise_for_status(st, info._err())
    if own:
        torch.cuda.current_stream(dev).synchronize()   # the plan's buffers die with `info`
        info.close()
