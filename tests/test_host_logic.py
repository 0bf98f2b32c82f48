"""Host-side argument decoding of the backend mirror (no GPU): view peeling, scaling
factors, conjugation rejection, shape errors with the reference's messages."""
import pytest
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import views


def _csr(m=3, n=4):
    return sb.csr_view(torch.zeros(4), torch.tensor([0, 3, 3, 4], dtype=torch.int32),
                       torch.tensor([2, 0, 2, 3], dtype=torch.int32), (m, n), 4)


def test_scaling_factor_product_and_base():
    a = _csr()
    x = torch.zeros(4)
    assert views.get_scaling_factor(a, x) is None
    assert views.get_scaling_factor(sb.scaled(2, a), sb.scaled(3, x)) == 6
    assert views.get_scaling_factor(sb.scaled(2, sb.matrix_opt(sb.scaled(5, a)))) == 10
    assert views.get_ultimate_base(sb.scaled(2, sb.matrix_opt(a))) is a
    assert views.is_conjugated(sb.conjugated(sb.scaled(2, a)))
    assert not views.is_conjugated(sb.conjugated(sb.conjugated(a)))


def test_transposed_reinterprets_arrays():
    a = _csr()
    t = sb.transposed(a)
    assert isinstance(t, sb.csc_view) and t.shape == (4, 3)
    assert t.colptr is a.rowptr and t.rowind is a.colind and t.values is a.values
    assert isinstance(sb.transposed(t), sb.csr_view)


def test_device_memory_required():
    a = _csr()
    with pytest.raises(RuntimeError, match="device memory"):
        sb.multiply(a, torch.zeros(4), torch.zeros(3))


def test_conjugated_rejected_like_cusparse_backend():
    a = _csr()
    with pytest.raises(RuntimeError, match="conjugated"):
        sb.multiply(sb.conjugated(a), torch.zeros(4), torch.zeros(3))


def test_bad_call_shapes():
    with pytest.raises(TypeError):
        sb.multiply(_csr())
    with pytest.raises(TypeError):
        sb.multiply_execute(None, _csr(), torch.zeros(4), torch.zeros(3))


def test_matrix_opt_detection():
    a = _csr()
    assert not views.has_matrix_opt(a) and not views.has_matrix_opt(sb.scaled(2, a))
    assert views.has_matrix_opt(sb.matrix_opt(a))
    assert views.has_matrix_opt(sb.scaled(2, sb.matrix_opt(sb.transposed(a))))


def test_transpose_argument_checks_follow_the_reference():
    # algorithms/transpose_impl.hpp:17-25: dimensions first, then the size of b's arrays
    a = _csr(3, 4)
    mk = lambda shape, cap: sb.csr_view(torch.zeros(cap), torch.zeros(shape[0] + 1, dtype=torch.int32),
                                        torch.zeros(cap, dtype=torch.int32), shape, 0)
    with pytest.raises(ValueError, match="transpose: matrix dimensions are incompatible."):
        sb.transpose(a, mk((3, 4), 4))
    with pytest.raises(RuntimeError, match="transpose: Transpose ran out of memory."):
        sb.transpose(a, mk((4, 3), 3))
    with pytest.raises(RuntimeError, match="device memory"):      # only then the backend's own rule
        sb.transpose(a, mk((4, 3), 4))
    with pytest.raises(TypeError):
        sb.transpose(a)
    with pytest.raises(TypeError):
        sb.transpose_inspect(sb.transposed(a), mk((4, 3), 4))      # csc_view: no overload either


def test_triangular_solve_argument_checks():
    sq = sb.csr_view(torch.zeros(3), torch.tensor([0, 1, 2, 3], dtype=torch.int32),
                     torch.tensor([0, 1, 2], dtype=torch.int32), (3, 3), 3)
    b, x = torch.zeros(3), torch.zeros(3)
    with pytest.raises(TypeError, match="uplo"):
        sb.triangular_solve(sq, "lower", sb.explicit_diagonal, b, x)
    with pytest.raises(TypeError, match="diag"):
        sb.triangular_solve(sq, sb.lower_triangle, None, b, x)
    with pytest.raises(ValueError, match="square"):
        sb.triangular_solve(_csr(3, 4), sb.lower_triangle, sb.explicit_diagonal, b, torch.zeros(4))
    with pytest.raises(ValueError, match="dimensions are incompatible"):
        sb.triangular_solve(sq, sb.upper_triangle, sb.implicit_unit_diagonal, torch.zeros(4), x)
    with pytest.raises(RuntimeError, match="conjugated"):
        sb.triangular_solve(sb.conjugated(sq), sb.lower_triangle, sb.explicit_diagonal, b, x)
    with pytest.raises(RuntimeError, match="device memory"):
        sb.triangular_solve(sq, sb.lower_triangle, sb.explicit_diagonal, sb.scaled(2.0, b), x)
    with pytest.raises(TypeError):
        sb.triangular_solve(sq, sb.lower_triangle, sb.explicit_diagonal, b)
    from spblas_reference_b200.triangular_solve import _own_scaling
    assert _own_scaling(sb.scaled(2.0, sb.matrix_opt(sb.scaled(3.0, sq)))) == 6.0
    assert _own_scaling(sq) is None


def test_host_execute_argument_checks():
    a = _csr()
    info = sb.operation_info_t()
    with pytest.raises(TypeError):
        sb.multiply_execute_host(None, a, torch.zeros(4), torch.zeros(3))
    with pytest.raises(RuntimeError, match="device memory"):      # A itself must be on the device
        sb.multiply_execute_host(info, a, torch.zeros(4), torch.zeros(3))


def test_dense_operand_generated_in_chunks_is_the_same_operand():
    """bench.py --workload c5mm fills its replicated B in row chunks (tens of GB must not
    triple their footprint while generated): same values as the one-shot generator."""
    from spblas_reference_b200 import generators as G
    for dt in (torch.float64, torch.float32, torch.int32):
        whole = G.dense_uniform((1000, 32), 6, dt, "cpu")
        parts = G.dense_uniform_rows(1000, 32, 6, dt, "cpu", chunk_elems=5000)
        assert torch.equal(whole, parts)


def test_hub_column_definition_known_answer(oracle):
    """oracle.hub_columns — the definition the GPU analysis (csrc/hub.cu) is compared with
    bit for bit: columns referenced >= min_count times, the max_cols most referenced (ties:
    smaller column first), renumbered ascending; a reference to hub s is re-encoded as ~s."""
    import numpy as np
    ci = np.array([5, 5, 5, 2, 2, 9, 9, 9, 9, 1, 0, 2], dtype=np.int32)
    hubs, refs, enc = oracle.hub_columns(ci, 10, 2, 2)      # counts: 9 -> 4, 2 -> 3, 5 -> 3
    assert hubs.tolist() == [2, 9] and refs == 7
    assert enc.tolist() == [5, 5, 5, -1, -1, -2, -2, -2, -2, 1, 0, -1]
    hubs, refs, enc = oracle.hub_columns(ci, 10, 8, 3)
    assert hubs.tolist() == [2, 5, 9] and refs == 10
    assert enc.tolist() == [-2, -2, -2, -1, -1, -3, -3, -3, -3, 1, 0, -1]
    hubs, refs, enc = oracle.hub_columns(ci, 10, 8, 5)      # nothing is referenced 5 times
    assert len(hubs) == 0 and refs == 0 and enc.tolist() == ci.tolist()
    hubs, refs, enc = oracle.hub_columns(np.array([], dtype=np.int32), 4, 8, 1)
    assert len(hubs) == 0 and refs == 0 and len(enc) == 0


def test_one_info_keeps_a_plan_per_inspected_structure(monkeypatch):
    """The reference's notes inspect `a` and `transposed(a)` with ONE operation_info_t and
    alternate the executes (notes/spmv.hpp:12-22): the info keeps one plan per structure
    (current + 3 parked, least recently used evicted) instead of re-inspecting at every
    switch.  Pure host logic: plan handles are faked."""
    import ctypes as C
    import sys
    M = sys.modules["spblas_reference_b200.multiply"]     # (the package exports the function)
    destroyed = []
    monkeypatch.setattr(M, "_destroy_plan", lambda plan: destroyed.append(plan.value))
    info = M.operation_info_t()
    assert not info._select("A")                      # nothing inspected yet

    def inspect(sig, handle):                         # what every _inspect does
        info._begin_inspect(sig)
        if not info._plan:
            info._plan = C.c_void_p(handle)           # (_ensure would create it)
        info._sig = sig
        return info._plan.value

    assert inspect("A", 1) == 1
    assert inspect("At", 2) == 2                      # A's plan is parked, not reused
    assert info._select("A") and info._plan.value == 1 and info._sig == "A"
    assert info._select("At") and info._plan.value == 2
    assert info._select("A") and info._plan.value == 1 and not destroyed
    assert inspect("A", 99) == 1                      # re-inspect: same plan, in place
    assert inspect("B", 3) == 3 and inspect("C", 4) == 4
    assert not destroyed                              # current C + parked At, A, B
    assert inspect("D", 5) == 5                       # a fifth structure evicts the oldest
    assert destroyed == [2] and not info._select("At")
    assert info._select("A") and info._plan.value == 1
    info.close()
    assert sorted(destroyed) == [1, 2, 3, 4, 5]
    assert not info._plan and info._parked == []
