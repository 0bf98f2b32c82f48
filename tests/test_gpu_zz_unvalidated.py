"""GPU tests of code written after this round's GPU time was spent: they have never run.
SPBLAS_B200_RUN_UNVALIDATED=1 runs them; each moves to its proper file with its first green
run on a B200.  The file name sorts last on purpose.

* the persistent, flag-synchronised triangular solve (csrc/trsv.cu: trsv_persistent_kernel,
  SPBLAS_B200_TRSV_PERSISTENT=1): same arithmetic per row as the level launches, so x must be
  BIT-IDENTICAL to the reference's, in ONE launch, and the give-up flag must stay 0;
* the hub-stream kernel with its memory gathers bypassing L1 (SPBLAS_B200_HUB_GATHER_CG=1):
  the same arithmetic, so y must be bit-identical to the warp-stream kernel's."""
import os
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from helpers import csr_on_device, dev
from test_gpu_trsv import DIAG, UPLO, _solve, _tri_matrix
from test_gpu_zhub import _lens, _run, _skewed_csr

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("SPBLAS_B200_RUN_UNVALIDATED") != "1",
                       reason="never run on a GPU yet (SPBLAS_B200_RUN_UNVALIDATED=1 runs it)"),
]


@pytest.mark.parametrize("ctas", ["0", "1"])
@pytest.mark.parametrize("kind", ["short", "mixed", "chain"])
@pytest.mark.parametrize("types", [(np.float32, np.int32, np.int32), (np.float64, np.int32, np.int64),
                                   (np.float32, np.int64, np.int64)])
def test_trsv_persistent_bit_exact(cuda, oracle, monkeypatch, kind, types, ctas):
    monkeypatch.setenv("SPBLAS_B200_TRSV_PERSISTENT", "1")
    monkeypatch.setenv("SPBLAS_B200_TRSV_CTAS_PER_SM", ctas)     # 1: a small grid, many rounds per thread
    vt, it, ot = types
    rng = np.random.default_rng(zlib.crc32(f"ptrsv{kind}{vt.__name__}{ot.__name__}".encode()))
    m = 40_000 if kind != "chain" else 3000                      # > one round of the 1-CTA-per-SM grid
    v, rp, ci = _tri_matrix(rng, m, kind, vt, it, ot)
    b = rng.standard_normal(m).astype(vt)
    a = csr_on_device(v, rp, ci, (m, m))
    for upper in (0, 1):
        for unit in (0, 1):
            info = sb.triangular_solve_inspect(a, UPLO[upper], DIAG[unit], dev(b),
                                               torch.empty(m, dtype=dev(b).dtype, device="cuda"))
            for kw in ({}, {"alpha_a": 0.5, "alpha_b": -2.0}):      # repeated solves: the epoch moves on
                want = oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit, **kw)
                got = _solve(a, upper, unit, b, m, info=info, **kw)
                assert np.array_equal(got, want, equal_nan=True), (kind, upper, unit, kw)
                assert info.last_launches == 1 and info.trsv_timeout == 0
            assert np.array_equal(_solve(a, upper, unit, b, m, info=info, inplace=True),
                                  oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit), equal_nan=True)
            assert info.trsv_timeout == 0
            info.close()


def test_trsv_persistent_poisson_wavefront(cuda, oracle, monkeypatch):
    """The stencil's anti-diagonal wavefront (1023 levels of <= 512 rows): every row waits on
    rows one level up, many levels are in flight at once."""
    monkeypatch.setenv("SPBLAS_B200_TRSV_PERSISTENT", "1")
    g = 512
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cuda:0")
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    m = shape[0]
    b = np.random.default_rng(3).standard_normal(m)
    want = None
    for tri in (sb.lower_triangle, sb.upper_triangle):
        x = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
        info = sb.triangular_solve_inspect(a, tri, sb.explicit_diagonal, dev(b), x)
        assert info.trsv_levels == 2 * g - 1
        for _ in range(3):
            sb.triangular_solve(info, a, tri, sb.explicit_diagonal, dev(b), x)
        torch.cuda.synchronize()
        assert info.last_launches == 1 and info.trsv_timeout == 0
        want = oracle.trsv(m, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy(), b,
                           upper=tri is sb.upper_triangle)
        assert np.array_equal(x.cpu().numpy(), want)
        info.close()


@pytest.mark.parametrize("kind", ["short", "hubrow", "long"])
@pytest.mark.parametrize("vt", [np.float32, np.float64, np.int32])
def test_hub_gathers_bypassing_l1(cuda, oracle, monkeypatch, kind, vt):
    rng = np.random.default_rng(zlib.crc32(f"hubcg{kind}{vt.__name__}".encode()))
    m, n = 5003, 2777
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, kind), vt)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    y_ws, i_ws = _run(a, xd, m, 2)
    monkeypatch.setenv("SPBLAS_B200_HUB_GATHER_CG", "1")          # read when the plan is created
    y_hub, i_hub = _run(a, xd, m, 3, hub=(64, 3))
    assert i_hub.spmv_variant == 3 and i_hub.hub_count == 64
    assert torch.equal(y_ws, y_hub)
    i_ws.close()
    i_hub.close()


@pytest.mark.parametrize("vt", [np.float32, np.float64])
def test_warp_stream_gathers_bypassing_l1(cuda, oracle, monkeypatch, vt):
    """SPBLAS_B200_WS_GATHER_CG=1: the plain walk with ld.global.cg gathers — the same sums."""
    rng = np.random.default_rng(77)
    m, n = 5003, 2777
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, "hubrow"), vt)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    y_ws, i_ws = _run(a, xd, m, 2)
    monkeypatch.setenv("SPBLAS_B200_WS_GATHER_CG", "1")
    y_cg, i_cg = _run(a, xd, m, 2)
    assert i_cg.spmv_variant == 2 and torch.equal(y_ws, y_cg)
    i_ws.close()
    i_cg.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_fuzz_spmv_kernels_on_arbitrary_small_structures(cuda, oracle, variant):
    """Property test (hypothesis): every SpMV kernel on arbitrary small structures — no rows,
    empty rows, one very long row, duplicates, a row block with a non-zero base — integer
    scalars, so the answer is exact whatever the order of the sums."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(0, 300), st.integers(1, 200), st.integers(0, 2 ** 31 - 1),
           st.sampled_from([0, 3, 9, 40]), st.integers(0, 700))
    def run(m, n, seed, maxlen, long_row):
        rng = np.random.default_rng(seed)
        lens = rng.integers(0, maxlen + 1, size=m)
        if m and long_row:
            lens[rng.integers(0, m)] = long_row
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(rp[-1])
        ci = rng.integers(0, n, size=nnz).astype(np.int32)
        v = rng.integers(-9, 10, size=nnz).astype(np.int32)
        x = rng.integers(-9, 10, size=n).astype(np.int32)
        a = csr_on_device(v, rp, ci, (m, n))
        xd = dev(x)
        y, info = _run(a, xd, m, variant, hub=(16, 1) if variant == 3 else None, alpha=2)
        assert np.array_equal(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=2))
        info.close()
        if m >= 4:                                   # rows [r0, r1) with the global base
            r0, r1 = m // 4, m - m // 4
            blk = sb.csr_view(a.values, a.rowptr[r0:r1 + 1], a.colind, (r1 - r0, n),
                              int(rp[r1] - rp[r0]))
            yb, ib = _run(blk, xd, r1 - r0, variant, hub=(16, 1) if variant == 3 else None)
            assert np.array_equal(yb.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v, x)[r0:r1])
            ib.close()

    run()
