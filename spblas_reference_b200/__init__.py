"""spblas_reference_b200 — B200 (sm_100a) backend for the sparse-times-dense hot path of
SparseBLAS/spblas-reference: multiply / multiply_inspect / multiply_execute on csr_view
and csc_view against a dense vector (SpMV) or a row-major matrix (SpMM).

The product is the C-ABI library libspblas_b200.so (hand-written CUDA, include/spblas_b200.h)
plus the C++20 backend headers under include/spblas/vendor/b200/.  This Python package is
the host-side mirror of the same interface used by the tests and the benchmark; it holds
no compute and no CPU fallback.
"""
from . import _cabi
from .multiply import (multiply, multiply_execute, multiply_execute_host, multiply_inspect,
                       operation_info_t)
from .transpose import transpose, transpose_inspect
from .triangular_solve import (explicit_diagonal, implicit_unit_diagonal, lower_triangle,
                               triangular_solve, triangular_solve_inspect, upper_triangle)
from .views import (conjugated, csc_view, csr_view, matrix_opt, scaled, scaled_view,
                    transposed)

__all__ = [
    "multiply", "multiply_inspect", "multiply_execute", "multiply_execute_host",
    "operation_info_t", "transpose", "transpose_inspect",
    "triangular_solve", "triangular_solve_inspect", "lower_triangle", "upper_triangle",
    "explicit_diagonal", "implicit_unit_diagonal",
    "csr_view", "csc_view", "scaled", "scaled_view", "transposed", "matrix_opt",
    "conjugated",
]
