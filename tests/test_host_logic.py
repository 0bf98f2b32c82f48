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
