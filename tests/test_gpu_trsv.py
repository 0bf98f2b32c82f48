"""GPU parity tests for triangular_solve_inspect / triangular_solve (level-scheduled SpTRSV,
csrc/trsv.cu): BIT-EXACT against the committed output of the real reference on its own
fixtures (test/gtest/triangular_solve_test.cpp: generate_csr over util::square_dims, values
scaled by 1e-3), against the oracle on every type combination, triangle, diagonal mode and
scaling, the level sets against their plain definition, and the error behaviour."""
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from conftest import GOLDEN
from helpers import csr_on_device, dev

pytestmark = pytest.mark.gpu
SQUARE_DIMS = [(1000, 1000, 100), (100, 100, 100), (40, 40, 1000)]
UPLO = {0: sb.lower_triangle, 1: sb.upper_triangle}
DIAG = {0: sb.explicit_diagonal, 1: sb.implicit_unit_diagonal}


def _solve(a, upper, unit, b, m, info=None, alpha_a=None, alpha_b=None, inplace=False):
    bd = dev(b)
    x = bd if inplace else torch.full((m,), float("nan"), dtype=bd.dtype, device="cuda")
    av = sb.scaled(alpha_a, a) if alpha_a is not None else a
    bv = sb.scaled(alpha_b, bd) if alpha_b is not None else bd
    if info is None:
        sb.triangular_solve(av, UPLO[upper], DIAG[unit], bv, x)
    else:
        sb.triangular_solve(info, av, UPLO[upper], DIAG[unit], bv, x)
    torch.cuda.synchronize()
    return x.cpu().numpy()


@pytest.mark.parametrize("dims", SQUARE_DIMS)
def test_reference_triangular_solve_fixtures(cuda, oracle, dims):
    g = np.load(f"{GOLDEN}/trsv_square_dims.npz")
    m, _, nnz = dims
    key = f"{m}_{nnz}"
    b = g[f"b_{key}"]
    a = csr_on_device(g[f"values_{key}"], g[f"ptr_{key}"], g[f"ind_{key}"], (m, m))
    ad = csr_on_device(g[f"dvalues_{key}"], g[f"dptr_{key}"], g[f"dind_{key}"], (m, m))
    for upper, name in ((0, "lower"), (1, "upper")):
        assert np.array_equal(_solve(a, upper, 1, b, m), g[f"x_unit_{name}_{key}"])
        assert np.array_equal(_solve(ad, upper, 0, b, m), g[f"x_explicit_{name}_{key}"])
        assert np.array_equal(_solve(ad, upper, 0, b, m, alpha_b=1.2),
                              g[f"x_explicit_scaled_{name}_{key}"])
        # the reference test itself (triangular_solve_test.cpp:70-88): b = 0, x starts at 1
        info = sb.triangular_solve_inspect(sb.matrix_opt(a), UPLO[upper], DIAG[1],
                                           dev(np.zeros(m, np.float32)), torch.ones(m, device="cuda"))
        x = torch.ones(m, device="cuda")
        sb.triangular_solve(info, sb.matrix_opt(a), UPLO[upper], DIAG[1], dev(np.zeros(m, np.float32)), x)
        assert not x.cpu().numpy().any()
        info.close()


def _tri_matrix(rng, m, kind, vt, it, ot):
    if kind == "short":
        lens = rng.integers(0, 10, size=m)
    elif kind == "mixed":
        lens = rng.integers(0, 5, size=m)
        lens[rng.integers(0, m, size=max(1, m // 40))] = rng.integers(40, 300, size=max(1, m // 40))
    else:                                          # "chain": every row reads its neighbour
        lens = np.ones(m, dtype=np.int64)
    rp = np.concatenate([[0], np.cumsum(lens + 1)]).astype(ot)          # +1: the diagonal
    ci = np.empty(int(rp[-1]), dtype=it)
    v = np.empty(int(rp[-1]), dtype=vt)
    for i in range(m):
        n_off = int(lens[i])
        cols = rng.integers(0, m, size=n_off) if kind != "chain" else np.array([max(i - 1, 0)])
        pos = int(rng.integers(0, n_off + 1))                           # diagonal anywhere in the row
        row_c = np.insert(cols, pos, i)
        row_v = np.insert(0.3 * rng.standard_normal(n_off) / max(n_off, 1), pos, 1.5 + rng.random())
        ci[rp[i]:rp[i + 1]] = row_c
        v[rp[i]:rp[i + 1]] = row_v
    return v, rp, ci


@pytest.mark.parametrize("graph", ["1", "0"])
@pytest.mark.parametrize("kind", ["short", "mixed", "chain"])
@pytest.mark.parametrize("types", [(np.float32, np.int32, np.int32), (np.float64, np.int32, np.int64),
                                   (np.float64, np.int32, np.int32), (np.float32, np.int64, np.int64)])
def test_trsv_bit_exact_vs_oracle(cuda, oracle, monkeypatch, kind, types, graph):
    # graph = 1: the level launches replayed from a CUDA graph (plans with >= 16 levels);
    # graph = 0: launched level by level
    monkeypatch.setenv("SPBLAS_B200_TRSV_GRAPH", graph)
    vt, it, ot = types
    rng = np.random.default_rng(zlib.crc32(f"trsv{kind}{vt.__name__}{ot.__name__}".encode()))
    m = 1500 if kind != "chain" else 700
    v, rp, ci = _tri_matrix(rng, m, kind, vt, it, ot)
    b = rng.standard_normal(m).astype(vt)
    a = csr_on_device(v, rp, ci, (m, m))
    for upper in (0, 1):
        for unit in (0, 1):
            info = sb.triangular_solve_inspect(a, UPLO[upper], DIAG[unit], dev(b),
                                               torch.empty(m, dtype=dev(b).dtype, device="cuda"))
            levels = oracle.trsv_levels(m, rp, ci, upper=bool(upper))
            assert info.trsv_levels == int(levels.max()) + 1
            for kw in ({}, {"alpha_b": 1.2}, {"alpha_a": 0.5, "alpha_b": -2.0}):
                want = oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit, **kw)
                got = _solve(a, upper, unit, b, m, info=info, **kw)
                assert np.array_equal(got, want, equal_nan=True), (kind, upper, unit, kw)
            assert info.last_launches == info.trsv_levels
            # no info (inspect + solve in one call), and b aliased with x
            assert np.array_equal(_solve(a, upper, unit, b, m), oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit),
                                  equal_nan=True)
            assert np.array_equal(_solve(a, upper, unit, b, m, info=info, inplace=True),
                                  oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit), equal_nan=True)
            # values change, structure does not
            v2 = v * vt(0.5)
            a.values.copy_(dev(v2))
            assert np.array_equal(_solve(a, upper, unit, b, m, info=info),
                                  oracle.trsv(m, rp, ci, v2, b, upper=upper, unit=unit), equal_nan=True)
            a.values.copy_(dev(v))
            info.close()


def test_trsv_poisson_levels_and_errors(cuda, oracle):
    g = 64
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cuda:0")
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    m = shape[0]
    b = np.random.default_rng(3).standard_normal(m)
    x = torch.empty(m, dtype=torch.float64, device="cuda")
    info = sb.triangular_solve_inspect(a, sb.lower_triangle, sb.explicit_diagonal, dev(b), x)
    assert info.trsv_levels == 2 * g - 1                    # anti-diagonals of the grid
    sb.triangular_solve(info, a, sb.lower_triangle, sb.explicit_diagonal, dev(b), x)
    want = oracle.trsv(m, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy(), b)
    assert np.array_equal(x.cpu().numpy(), want)
    info.close()
    # a row without a stored diagonal under explicit_diagonal: refused at inspect
    rp2 = np.array([0, 1, 2, 3], np.int32)
    a2 = csr_on_device(np.ones(3, np.float32), rp2, np.array([0, 0, 2], np.int32), (3, 3))
    bb, xx = torch.ones(3, device="cuda"), torch.empty(3, device="cuda")
    with pytest.raises(RuntimeError, match="no diagonal"):
        sb.triangular_solve(a2, sb.lower_triangle, sb.explicit_diagonal, bb, xx)
    sb.triangular_solve(a2, sb.lower_triangle, sb.implicit_unit_diagonal, bb, xx)   # fine without
    assert xx.cpu().tolist() == [1.0, 0.0, 1.0]
    with pytest.raises(ValueError):                          # not square
        sb.triangular_solve(csr_on_device(np.ones(1, np.float32), np.array([0, 1, 1], np.int32),
                                          np.array([0], np.int32), (2, 3)),
                            sb.lower_triangle, sb.implicit_unit_diagonal, torch.ones(2, device="cuda"),
                            torch.ones(3, device="cuda"))
    with pytest.raises(ValueError):                          # wrong vector length
        sb.triangular_solve(a2, sb.lower_triangle, sb.implicit_unit_diagonal,
                            torch.ones(4, device="cuda"), xx)


@pytest.mark.parametrize("inspect", ["frontier", "relax"])
def test_trsv_level_analysis_variants(cuda, oracle, monkeypatch, inspect):
    """The frontier (Kahn, default) and the relaxation-sweep level analyses find the same
    level count and give the same bit-exact solution; duplicate dependency entries and
    rows with no entries at all included."""
    monkeypatch.setenv("SPBLAS_B200_TRSV_INSPECT", inspect)
    rng = np.random.default_rng(77)
    m = 2000
    v, rp, ci = _tri_matrix(rng, m, "mixed", np.float64, np.int32, np.int32)
    # duplicate some dependency entries; the matrix keeps its diagonals
    ci[rp[500] + 1:rp[501]] = ci[rp[500] + 1] if rp[501] - rp[500] > 1 else ci[rp[500]]
    b = rng.standard_normal(m)
    a = csr_on_device(v, rp, ci, (m, m))
    for upper in (0, 1):
        info = sb.triangular_solve_inspect(a, UPLO[upper], DIAG[0], dev(b),
                                           torch.empty(m, dtype=torch.float64, device="cuda"))
        assert info.trsv_levels == int(oracle.trsv_levels(m, rp, ci, upper=bool(upper)).max()) + 1
        assert np.array_equal(_solve(a, upper, 0, b, m, info=info),
                              oracle.trsv(m, rp, ci, v, b, upper=upper, unit=False), equal_nan=True)
        info.close()


def test_c_abi_guards_found_by_review(cuda, oracle):
    """Round-1 review: (1) spblas_b200_inspect on a plan that holds a triangular
    structure must invalidate it — the two share index types and owned buffers, and a later
    spblas_b200_trsv would read the triangle with the new widths; (2) trsv_inspect must not
    trust the caller's nnz or an offsets array that is not monotone."""
    import ctypes as C
    from spblas_reference_b200 import _cabi
    L = _cabi.lib()
    rng = np.random.default_rng(8)
    m = 500
    v, rp, ci = _tri_matrix(rng, m, "short", np.float64, np.int64, np.int64)
    a64 = csr_on_device(v, rp, ci, (m, m))
    b = dev(rng.standard_normal(m))
    x = torch.empty(m, dtype=torch.float64, device="cuda")
    info = sb.triangular_solve_inspect(a64, sb.lower_triangle, sb.implicit_unit_diagonal, b, x)
    # the same plan now inspects an int32 matrix for a product ...
    a32 = csr_on_device(v.astype(np.float64), rp.astype(np.int32), ci.astype(np.int32), (m, m))
    st = L.spblas_b200_inspect(info._plan, _cabi.CSR, m, m, a32.nnz, a32.rowptr.data_ptr(),
                               a32.colind.data_ptr(), _cabi.I32, _cabi.I32, 1, 0)
    assert st == _cabi.SUCCESS
    # ... and the triangular solve must refuse instead of walking int64 arrays as int32
    st = L.spblas_b200_trsv(info._plan, _cabi.F64, None, None, a64.values.data_ptr(),
                            b.data_ptr(), x.data_ptr())
    assert st == _cabi.NOT_INSPECTED
    info.close()

    plan = C.c_void_p()
    assert L.spblas_b200_plan_create(C.byref(plan), None) == 0
    try:
        a = csr_on_device(v, rp.astype(np.int32), ci.astype(np.int32), (m, m))
        st = L.spblas_b200_trsv_inspect(plan, m, a.nnz + 7, a.rowptr.data_ptr(),
                                        a.colind.data_ptr(), _cabi.I32, _cabi.I32, 0, 1)
        assert st == _cabi.INVALID_STRUCTURE, L.spblas_b200_last_error(plan)
        bad = rp.astype(np.int32).copy()
        bad[10], bad[11] = bad[11] + 3, bad[10]
        badp = dev(bad)
        st = L.spblas_b200_trsv_inspect(plan, m, a.nnz, badp.data_ptr(), a.colind.data_ptr(),
                                        _cabi.I32, _cabi.I32, 0, 1)
        assert st == _cabi.INVALID_STRUCTURE
        st = L.spblas_b200_trsv_inspect(plan, m, a.nnz, a.rowptr.data_ptr(), a.colind.data_ptr(),
                                        _cabi.I32, _cabi.I32, 0, 1)
        assert st == _cabi.SUCCESS
    finally:
        L.spblas_b200_plan_destroy(plan)


def test_csc_index_outside_the_matrix_is_refused(cuda):
    """A row index >= m in a CSC operand: the image builder's sort looks at the low bits only,
    so the inspect phase must check the range itself."""
    from helpers import csc_on_device
    colptr = np.array([0, 2, 3, 5], np.int32)
    rowind = np.array([0, 9, 1, 2, 3], np.int32)          # 9 >= m = 4
    a = csc_on_device(np.ones(5, np.float32), colptr, rowind, (4, 3))
    with pytest.raises(RuntimeError, match="outside the matrix"):
        sb.multiply_inspect(a, torch.ones(3, device="cuda"), torch.empty(4, device="cuda"))
