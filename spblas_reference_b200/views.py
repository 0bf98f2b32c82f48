"""Host-side mirror of the reference's views for the sparse-times-dense path.

Same names and meaning as the reference's C++ types; storage is torch CUDA tensors
(device memory only — PyTorch is plumbing here, the compute is the C-ABI library):

  csr_view(values, rowptr, colind, shape, nnz)   reference views/csr_view.hpp:12-77
  csc_view(values, colptr, rowind, shape, nnz)   reference views/csc_view.hpp:9-72
  scaled(alpha, t)                               reference algorithms/scaled_impl.hpp:8-16,
                                                 views/scaled_view_impl.hpp:20-91,97-219
  transposed(a)                                  reference algorithms/transposed.hpp:7-21
  matrix_opt(a)                                  reference views/matrix_opt_impl.hpp:14-93
  conjugated(t)                                  reference views/conjugated_view_impl.hpp (rejected
                                                 by GPU backends: vendor/cusparse/spmv_impl.hpp:32-36)

Views own nothing; the tensors must outlive the calls that use them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Tuple

import torch

from . import _cabi

_VAL = {torch.float32: _cabi.F32, torch.float64: _cabi.F64, torch.int32: _cabi.S32}
_IDX = {torch.int32: _cabi.I32, torch.int64: _cabi.I64}


def value_type(t: torch.Tensor) -> int:
    try:
        return _VAL[t.dtype]
    except KeyError:
        # complex / half etc.: the reference's GPU type gate
        # (vendor/cusparse/types.hpp:17-20) has no overload either
        raise RuntimeError(f"b200 backend does not support scalar type {t.dtype}") from None


def index_type(t: torch.Tensor) -> int:
    try:
        return _IDX[t.dtype]
    except KeyError:
        raise RuntimeError(f"b200 backend needs int32/int64 indices, got {t.dtype}") from None


def _check_1d_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor holding device memory")
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live in device memory (the b200 backend has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")
    return t


@dataclass
class csr_view:
    values: torch.Tensor
    rowptr: torch.Tensor
    colind: torch.Tensor
    shape: Tuple[int, int]
    nnz: int

    def __post_init__(self):
        self.shape = (int(self.shape[0]), int(self.shape[1]))
        self.nnz = int(self.nnz)

    def size(self) -> int:
        return self.nnz


@dataclass
class csc_view:
    values: torch.Tensor
    colptr: torch.Tensor
    rowind: torch.Tensor
    shape: Tuple[int, int]
    nnz: int

    def __post_init__(self):
        self.shape = (int(self.shape[0]), int(self.shape[1]))
        self.nnz = int(self.nnz)

    def size(self) -> int:
        return self.nnz


@dataclass
class scaled_view:
    alpha: Any
    base: Any


@dataclass
class conjugated_view:
    base: Any


@dataclass
class matrix_opt:
    """Transparent wrapper; the reference uses it to carry a vendor handle
    (views/matrix_opt_impl.hpp:88-92) and its oneMKL backend optimises only operands
    wrapped in it (vendor/onemkl_sycl/spmv_impl.hpp:46-58).  Here the plan lives in
    operation_info_t; multiply_inspect on a matrix_opt additionally lets the plan keep
    value-dependent state (the values of a CSC / transposed operand gathered into image
    order): the values must then stay unchanged until the next multiply_inspect."""
    base: Any


def scaled(alpha, t):
    return scaled_view(alpha, t)


def conjugated(t):
    return conjugated_view(t)


def transposed(a):
    """transposed(csr) is the csc over the same arrays and vice versa
    (reference algorithms/transposed.hpp:7-21)."""
    if isinstance(a, csr_view):
        return csc_view(a.values, a.rowptr, a.colind, (a.shape[1], a.shape[0]), a.nnz)
    if isinstance(a, csc_view):
        return csr_view(a.values, a.colptr, a.rowind, (a.shape[1], a.shape[0]), a.nnz)
    raise TypeError("transposed() expects a csr_view or csc_view")


# ---- view introspection (reference detail/view_inspectors.hpp) ---------------------
def get_ultimate_base(t):
    """detail/view_inspectors.hpp:104-111"""
    while isinstance(t, (scaled_view, conjugated_view, matrix_opt)):
        t = t.base
    return t


def get_scaling_factor(*tensors) -> Optional[Any]:
    """Product of all scaled_view factors in the chains, or None
    (detail/view_inspectors.hpp:22-77)."""
    out = None
    for t in tensors:
        while isinstance(t, (scaled_view, conjugated_view, matrix_opt)):
            if isinstance(t, scaled_view):
                out = t.alpha if out is None else out * t.alpha
            t = t.base
    return out


def has_matrix_opt(t) -> bool:
    """detail/view_inspectors.hpp:113-122"""
    while isinstance(t, (scaled_view, conjugated_view, matrix_opt)):
        if isinstance(t, matrix_opt):
            return True
        t = t.base
    return False


def is_conjugated(t) -> bool:
    """detail/view_inspectors.hpp:81-97 (odd number of conjugated views)."""
    c = False
    while isinstance(t, (scaled_view, conjugated_view, matrix_opt)):
        if isinstance(t, conjugated_view):
            c = not c
        t = t.base
    return c
