"""The 4-argument multiply, y = alpha A x + beta d and C = alpha A B + beta D (SURVEY 8f n3;
reference convention: vendor/rocsparse/multiply_spgemm.hpp:69-118 — beta is the scaling factor
of d), fused into every kernel's single store per row: every SpMV variant, both SpMM kernels,
split rows, in place (d is y), beta = 0 with a poisoned d, the no-info overloads — and the
structure cache behind those overloads (a structure seen before is reused after a device-side
check of its offsets array; a change in place, or a malformed array, must be noticed)."""
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from helpers import EPS, csc_on_device, csr_on_device, dev
from test_gpu_zhub import _lens, _skewed_csr

pytestmark = pytest.mark.gpu


def _check(got, want, rp, bound, what):
    """|got - want| <= (len_i + 3) eps (sum |alpha a x| + |beta d|)."""
    got, want = np.asarray(got), np.asarray(want)
    if want.dtype.kind != "f":
        assert np.array_equal(got, want), what
        return
    lens = np.diff(np.asarray(rp).astype(np.int64)).astype(np.float64)
    if want.ndim == 2:
        lens = lens[:, None]
    tol = (lens + 3.0) * EPS[want.dtype] * bound
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    assert np.isfinite(got).all(), what
    assert (err <= tol).all(), (what, float((err / np.maximum(tol, 1e-300)).max()))


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("vt", [np.float32, np.float64, np.int32])
def test_spmv_axpby_every_kernel(cuda, oracle, variant, vt):
    rng = np.random.default_rng(zlib.crc32(f"axpby{variant}{vt.__name__}".encode()))
    m, n = 5003, 2777
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, "hubrow"), vt)
    d = (rng.integers(-9, 10, size=m) if vt is np.int32 else rng.standard_normal(m)).astype(vt)
    a = csr_on_device(v, rp, ci, (m, n))
    xd, dd = dev(x), dev(d)
    alpha, beta = (3, -2) if vt is np.int32 else (0.75, -1.5)
    y = torch.full((m,), float("nan") if vt is not np.int32 else 77, dtype=xd.dtype, device="cuda")
    info = sb.multiply_inspect(a, xd, y)
    if variant >= 3:
        info.set_hub(True, 64, 3)
    info.force_spmv_variant(variant)
    sb.multiply(info, sb.scaled(alpha, a), xd, y, sb.scaled(beta, dd))
    assert info.spmv_variant == variant
    t = oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=alpha)
    want = oracle.axpby(t, d, beta)
    bound = None if vt is np.int32 else oracle.abs_rowsum(rp, ci, v, x, alpha) + np.abs(beta * d.astype(np.float64))
    _check(y.cpu().numpy(), want, rp, bound, f"variant {variant}")
    # beta = 1 (d not scaled), in place: y <- alpha A x + y
    y.copy_(dd)
    sb.multiply_execute(info, sb.scaled(alpha, a), xd, y, y)
    want1 = oracle.axpby(t, d, 1)
    bound1 = None if vt is np.int32 else oracle.abs_rowsum(rp, ci, v, x, alpha) + np.abs(d.astype(np.float64))
    _check(y.cpu().numpy(), want1, rp, bound1, f"variant {variant} in place")
    # beta = 0: d is not read — a NaN in it must not reach y
    if vt is not np.int32:
        poison = torch.full_like(dd, float("nan"))
        sb.multiply(info, sb.scaled(alpha, a), xd, y, sb.scaled(0.0, poison))
        _check(y.cpu().numpy(), t, rp, oracle.abs_rowsum(rp, ci, v, x, alpha), "beta = 0")
    info.close()


def test_spmv_axpby_csc_and_no_info(cuda, oracle):
    rng = np.random.default_rng(11)
    m, n = 1200, 900
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, "short"), np.float64)
    # the same matrix column-major
    import scipy.sparse as sp
    csc = sp.csr_matrix((v, ci, rp), shape=(m, n)).tocsc()
    ac = csc_on_device(csc.data, csc.indptr.astype(np.int32), csc.indices.astype(np.int32), (m, n))
    d = rng.standard_normal(m)
    xd, dd = dev(x), dev(d)
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    sb.multiply(ac, xd, y, sb.scaled(2.0, dd))                     # no info, CSC
    t = oracle.spmv("csc", (m, n), csc.indptr.astype(np.int32), csc.indices.astype(np.int32), csc.data, x)
    bound = oracle.abs_rowsum(rp, ci, v, x) + np.abs(2.0 * d)
    _check(y.cpu().numpy(), oracle.axpby(t, d, 2.0), rp, bound, "csc, no info")
    a = csr_on_device(v, rp, ci, (m, n))
    for _ in range(3):                                              # no info, CSR: cached structure
        y.fill_(float("nan"))
        sb.multiply(a, xd, y, sb.scaled(2.0, dd))
        _check(y.cpu().numpy(), oracle.axpby(oracle.spmv("csr", (m, n), rp, ci, v, x), d, 2.0), rp,
               bound, "csr, no info")
    with pytest.raises(ValueError, match="dimensions are incompatible"):
        sb.multiply(a, xd, y, torch.zeros(m + 1, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("k", [1, 8, 32, 128, 200])
@pytest.mark.parametrize("vt", [np.float32, np.float64])
def test_spmm_axpby(cuda, oracle, k, vt):
    rng = np.random.default_rng(zlib.crc32(f"mmaxpby{k}{vt.__name__}".encode()))
    m, n = 900, 700
    lens = _lens(rng, m, "short")
    lens[17] = 6000                                     # one row cut into segments
    v, rp, ci, _ = _skewed_csr(rng, m, n, lens, vt)
    B = rng.standard_normal((n, k)).astype(vt)
    D = rng.standard_normal((m, k)).astype(vt)
    a = csr_on_device(v, rp, ci, (m, n))
    Bd, Dd = dev(B), dev(D)
    alpha, beta = 0.5, -2.0
    t = oracle.spmm("csr", (m, n), rp, ci, v, B, alpha_a=alpha)
    want = oracle.axpby(t, D, beta)
    from helpers import spmm_bound
    bound = spmm_bound(rp, ci, v, B, alpha) + np.abs(beta * D.astype(np.float64))
    for forced in ("0", "1"):                            # row kernel, stream (ring) kernel
        import os
        os.environ["SPBLAS_B200_SPMM_VARIANT"] = forced
        try:
            C = torch.full((m, k), float("nan"), dtype=Bd.dtype, device="cuda")
            info = sb.multiply_inspect(a, Bd, C)
            sb.multiply(info, sb.scaled(alpha, a), Bd, C, sb.scaled(beta, Dd))
            _check(C.cpu().numpy(), want, rp, bound, f"spmm k={k} forced={forced}")
            C.copy_(Dd)                                   # in place: C <- alpha A B + beta C
            sb.multiply_execute(info, sb.scaled(alpha, a), Bd, C, sb.scaled(beta, C))
            _check(C.cpu().numpy(), want, rp, bound, f"spmm k={k} forced={forced} in place")
            info.close()
        finally:
            os.environ.pop("SPBLAS_B200_SPMM_VARIANT")
    C = torch.empty((m, k), dtype=Bd.dtype, device="cuda")
    sb.multiply(sb.scaled(alpha, a), Bd, C, sb.scaled(beta, Dd))    # no info
    _check(C.cpu().numpy(), want, rp, bound, f"spmm k={k} no info")


# ---- the structure cache of the no-info overloads ------------------------------------------
def test_no_info_overload_notices_a_structure_changed_in_place(cuda, oracle):
    rng = np.random.default_rng(21)
    m, n = 3000, 2500
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, "short"), np.float32)
    nnz = int(rp[-1])
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    y = torch.empty(m, device="cuda")
    for _ in range(3):                                   # first call inspects, the next two reuse
        y.fill_(float("nan"))
        sb.multiply(a, xd, y)
        _check(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v, x), rp,
               oracle.abs_rowsum(rp, ci, v, x), "cached structure")
    # the same arrays, another structure with the same nnz: rows re-cut
    lens2 = np.diff(rp).copy()
    rng.shuffle(lens2)
    rp2 = np.concatenate([[0], np.cumsum(lens2)]).astype(np.int32)
    assert int(rp2[-1]) == nnz
    a.rowptr.copy_(dev(rp2))
    for _ in range(2):
        y.fill_(float("nan"))
        sb.multiply(a, xd, y)
        _check(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp2, ci, v, x), rp2,
               oracle.abs_rowsum(rp2, ci, v, x), "structure changed in place")
    # malformed in place: must raise, on the cached path too
    bad = rp2.copy()
    bad[5], bad[6] = bad[6] + 1, bad[5]
    a.rowptr.copy_(dev(bad))
    with pytest.raises(RuntimeError, match="monoton|span"):
        sb.multiply(a, xd, y)
    a.rowptr.copy_(dev(rp2))                             # and recovers
    sb.multiply(a, xd, y)
    _check(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp2, ci, v, x), rp2,
           oracle.abs_rowsum(rp2, ci, v, x), "after the malformed call")


def test_no_info_overload_alternating_matrices_and_streams(cuda, oracle):
    rng = np.random.default_rng(22)
    mats = []
    for i, kind in enumerate(["short", "hubrow"]):
        m, n = 2000 + 100 * i, 1500
        v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, kind), np.float64)
        mats.append((csr_on_device(v, rp, ci, (m, n)), dev(x), v, rp, ci, x, m, n))
    side = torch.cuda.Stream()
    for rep in range(4):
        for a, xd, v, rp, ci, x, m, n in mats:
            y = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
            if rep % 2:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    sb.multiply(a, xd, y)
                torch.cuda.current_stream().wait_stream(side)
            else:
                sb.multiply(a, xd, y)
            _check(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v, x), rp,
                   oracle.abs_rowsum(rp, ci, v, x), f"rep {rep}")
