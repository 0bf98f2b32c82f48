"""GPU parity tests for transpose_inspect / transpose (CSR -> CSR on the device): bit-exact
against the committed output of the real reference on its own fixtures
(test/gtest/transpose_test.cpp), against the oracle on every type combination and row-length
mix, the reference's error behaviour, and the round trip transpose(transpose(A))."""
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from conftest import DIMS, golden
from helpers import csr_on_device, dev
from test_gpu_spmv import _random_csr

pytestmark = pytest.mark.gpu


def _empty_b(a, m, n, slack=0):
    nnz = a.nnz
    return sb.csr_view(torch.full((nnz + slack,), 7, dtype=a.values.dtype, device="cuda"),
                       torch.full((n + 1,), -5, dtype=a.rowptr.dtype, device="cuda"),
                       torch.full((nnz + slack,), -5, dtype=a.colind.dtype, device="cuda"),
                       (n, m), 0)


def _host(b):
    return (b.values[:b.nnz].cpu().numpy(), b.rowptr.cpu().numpy(), b.colind[:b.nnz].cpu().numpy())


@pytest.mark.parametrize("dims", DIMS)
def test_reference_transpose_test(cuda, dims):
    g = golden(*dims)
    m, n, _ = dims
    a = csr_on_device(g["csr_values"], g["csr_ptr"], g["csr_ind"], (m, n))
    b = _empty_b(a, m, n)
    info = sb.transpose_inspect(a, b)                    # transpose_test.cpp:33-34
    sb.transpose(info, a, b)
    tv, trp, tci = _host(b)
    assert b.nnz == a.nnz
    assert np.array_equal(tv, g["csr_transpose_values"])
    assert np.array_equal(trp, g["csr_transpose_ptr"])
    assert np.array_equal(tci, g["csr_transpose_ind"])


@pytest.mark.parametrize("kind", ["short", "mixed", "hub", "empty"])
@pytest.mark.parametrize("types", [(np.float32, np.int32, np.int32), (np.float64, np.int32, np.int64),
                                   (np.int32, np.int32, np.int32), (np.float32, np.int64, np.int64)])
def test_transpose_vs_oracle(cuda, oracle, kind, types):
    vt, it, ot = types
    rng = np.random.default_rng(zlib.crc32(f"tr{kind}{vt.__name__}{ot.__name__}".encode()))
    m, n = 3001, 1777
    v, rp, ci, _ = _random_csr(rng, m, n, kind, vt, it, ot)
    a = csr_on_device(v, rp, ci, (m, n))
    b = _empty_b(a, m, n, slack=3)                       # larger arrays are fine
    sb.transpose(a, b)                                   # the overload without info
    want = oracle.transpose((m, n), rp, ci, v)
    for got, w in zip(_host(b), want):
        assert got.dtype == w.dtype and np.array_equal(got, w)
    # values change, structure does not: the inspected plan is reused
    info = sb.transpose_inspect(a, b)
    for rep in range(2):
        v2 = (rng.integers(-99, 99, size=len(v)) if vt == np.int32 else rng.standard_normal(len(v))).astype(vt)
        a.values.copy_(dev(v2))
        sb.transpose(info, a, b)
        want = oracle.transpose((m, n), rp, ci, v2)
        for got, w in zip(_host(b), want):
            assert np.array_equal(got, w)
    # the same plan multiplies: it is the plan of transposed(a)
    if vt != np.int32 and len(v):
        x = rng.standard_normal(m).astype(vt)
        y = torch.empty(n, dtype=a.values.dtype, device="cuda")
        sb.multiply(sb.transposed(a), dev(x), y)
        yb = torch.empty_like(y)
        sb.multiply(b, dev(x), yb)                       # B = A^T as a plain CSR matrix
        torch.cuda.synchronize()
        assert np.allclose(y.cpu().numpy(), yb.cpu().numpy(), rtol=1e-4 if vt == np.float32 else 1e-12,
                           atol=1e-3 if vt == np.float32 else 1e-10)
    info.close()


def test_transpose_round_trip_and_errors(cuda, oracle):
    rng = np.random.default_rng(21)
    m, n = 900, 1300
    v, rp, ci, _ = _random_csr(rng, m, n, "short", np.float64, np.int32, np.int32)
    a = csr_on_device(v, rp, ci, (m, n))
    b = _empty_b(a, m, n)
    sb.transpose(a, b)
    c = _empty_b(b, n, m)
    sb.transpose(b, c)
    # (A^T)^T has A's rows with the entries of every row in ascending column order (stable)
    cv, crp, cci = _host(c)
    assert np.array_equal(crp, rp)
    for i in range(0, m, 37):
        s = slice(rp[i], rp[i + 1])
        order = np.argsort(ci[s], kind="stable")
        assert np.array_equal(cci[s], ci[s][order]) and np.array_equal(cv[s], v[s][order])
    with pytest.raises(ValueError, match="dimensions are incompatible"):
        sb.transpose(a, _empty_b(a, m, n + 1))
    small = sb.csr_view(torch.empty(max(a.nnz - 1, 0), dtype=torch.float64, device="cuda"),
                        torch.empty(n + 1, dtype=torch.int32, device="cuda"),
                        torch.empty(max(a.nnz - 1, 0), dtype=torch.int32, device="cuda"), (n, m), 0)
    with pytest.raises(RuntimeError, match="ran out of memory"):
        sb.transpose(a, small)
