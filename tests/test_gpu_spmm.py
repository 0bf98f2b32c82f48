"""GPU parity tests for SpMM (row-major B and C).  Re-hosts test/gtest/spmm_test.cpp:6-221
(CsrView SpMM / SpMM_AScaled / SpMM_BScaled / SpMM_Aopt, CscView SpMM) on the device — the
reference itself has no device SpMM test — and adds wide/narrow/odd widths, padded leading
dimensions, fp64, int32 scalars, int64 offsets and split (hub) rows."""
import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from conftest import DIMS, golden
from helpers import (assert_rows_within_bound, csc_on_device, csr_on_device, dev, gpu_spmm,
                     spmm_bound)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims", DIMS)
@pytest.mark.parametrize("k", [1, 8, 32, 64, 512])
def test_reference_spmm_tests(cuda, oracle, dims, k):
    g = golden(*dims)
    m, n, _ = dims
    v, rp, ci = g["csr_values"], g["csr_ptr"], g["csr_ind"]
    B = g[f"dense_B_{k}"]
    a = csr_on_device(v, rp, ci, (m, n))
    C = gpu_spmm(a, B, m)
    assert oracle.expect_eq_tolerance(g[f"csr_spmm_{k}"], C).all()
    assert_rows_within_bound(C, g[f"csr_spmm_{k}"], rp, spmm_bound(rp, ci, v, B), "spmm")
    if k <= 64:
        Ca = gpu_spmm(a, B, m, alpha_a=2.0)                      # SpMM_AScaled
        assert oracle.expect_eq_tolerance(g[f"csr_spmm_ascaled_{k}"], Ca).all()
        Cb = gpu_spmm(a, B, m, alpha_b=2.0)                      # SpMM_BScaled
        assert oracle.expect_eq_tolerance(oracle.spmm("csr", (m, n), rp, ci, v, B, alpha_b=2.0),
                                          Cb).all()
        Co = gpu_spmm(sb.matrix_opt(a), B, m)                    # SpMM_Aopt
        assert np.array_equal(Co, C)
    # CscView.SpMM
    vc, cp, ri = g["csc_values"], g["csc_ptr"], g["csc_ind"]
    ac = csc_on_device(vc, cp, ri, (m, n))
    Cc = gpu_spmm(ac, B, m)
    assert oracle.expect_eq_tolerance(g[f"csc_spmm_{k}"], Cc).all()


@pytest.mark.parametrize("k", [1, 3, 4, 7, 16, 33, 128, 130])
@pytest.mark.parametrize("types", [(np.float32, np.int32, np.int32),
                                   (np.float64, np.int32, np.int32),
                                   (np.int32, np.int32, np.int32),
                                   (np.float64, np.int32, np.int64)])
def test_spmm_widths_and_types(cuda, oracle, k, types):
    vt, it, ot = types
    rng = np.random.default_rng(k * 17 + 1)
    m, n = 1203, 801
    lens = rng.integers(0, 24, size=m)
    lens[7] = 300
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(ot)
    ci = rng.integers(0, n, size=int(rp[-1])).astype(it)
    if vt == np.int32:
        v = rng.integers(-5, 6, size=len(ci)).astype(vt)
        B = rng.integers(-5, 6, size=(n, k)).astype(vt)
        alpha = 2
    else:
        v = rng.standard_normal(len(ci)).astype(vt)
        B = rng.standard_normal((n, k)).astype(vt)
        alpha = -1.5
    a = csr_on_device(v, rp, ci, (m, n))
    for kw in ({}, {"alpha_a": alpha}):
        C_ref = oracle.spmm("csr", (m, n), rp, ci, v, B, **kw)
        bound = None if vt == np.int32 else spmm_bound(rp, ci, v, B, kw.get("alpha_a", 1.0))
        Bd = dev(B)
        info = sb.multiply_inspect(a, Bd, torch.empty((m, k), dtype=Bd.dtype, device="cuda"))
        for C in (gpu_spmm(a, B, m, **kw), gpu_spmm(a, B, m, info=info, **kw)):
            assert_rows_within_bound(C, C_ref, rp, bound, f"k={k} {vt.__name__}")
        info.close()


def test_spmm_padded_leading_dimensions_and_nan_isolation(cuda, oracle):
    rng = np.random.default_rng(9)
    m, n, k, ld = 500, 400, 32, 40
    lens = rng.integers(0, 20, size=m)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = rng.integers(1, n, size=int(rp[-1])).astype(np.int32)   # row 0 of B never referenced
    v = rng.standard_normal(len(ci)).astype(np.float32)
    B = rng.standard_normal((n, k)).astype(np.float32)
    B[0] = np.nan                                                # must not leak (0 * NaN)
    a = csr_on_device(v, rp, ci, (m, n))
    Bp = torch.full((n, ld), float("nan"), device="cuda")
    Bp[:, :k] = dev(B)
    Cp = torch.full((m, ld), -7.0, device="cuda")
    sb.multiply(a, Bp[:, :k], Cp[:, :k])                         # strided views: ld = 40
    C = Cp[:, :k].cpu().numpy()
    B0 = B.copy()
    B0[0] = 0
    assert_rows_within_bound(C, oracle.spmm("csr", (m, n), rp, ci, v, B0), rp,
                             spmm_bound(rp, ci, v, B0), "padded")
    assert (Cp[:, k:] == -7.0).all()                             # padding untouched
    with pytest.raises(ValueError, match="matrix dimensions are incompatible"):
        sb.multiply(a, Bp[:, :k], torch.empty((m, k + 1), device="cuda"))


def test_spmm_split_hub_rows(cuda, oracle):
    """Rows longer than the segment limit are cut into segments and recombined."""
    rng = np.random.default_rng(13)
    m, n, k = 300, 5000, 32
    lens = rng.integers(0, 8, size=m)
    lens[3] = 10000
    lens[299] = 4097
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = rng.integers(0, n, size=int(rp[-1])).astype(np.int32)
    v = rng.standard_normal(len(ci))
    B = rng.standard_normal((n, k))
    a = csr_on_device(v, rp, ci, (m, n))
    Bd = dev(B)
    Cd = torch.empty((m, k), dtype=torch.float64, device="cuda")
    info = sb.multiply_inspect(a, Bd, Cd)
    assert info.num_segments == 3 + 2
    sb.multiply(info, a, Bd, Cd)
    assert_rows_within_bound(Cd.cpu().numpy(), oracle.spmm("csr", (m, n), rp, ci, v, B), rp,
                             spmm_bound(rp, ci, v, B), "hub spmm")
    # the no-info overload takes the same path
    C2 = gpu_spmm(a, B, m)
    assert np.array_equal(C2, Cd.cpu().numpy())


@pytest.mark.parametrize("k", [32, 128])
def test_c3_shape_reduced(cuda, oracle, k):
    m = n = 200_000                                              # 2M x 2M runs in bench.py
    v, rp, ci, shape = G.uniform_random_csr(m, n, 16, seed=3, dtype=torch.float32, device=cuda)
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    B = G.dense_uniform((n, k), 4, torch.float32, cuda)
    C = torch.empty((m, k), device=cuda)
    info = sb.multiply_inspect(a, B, C)
    sb.multiply(info, a, B, C)
    vh, rph, cih, Bh = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy(), B.cpu().numpy()
    C_ref = oracle.spmm("csr", shape, rph, cih, vh, Bh)
    # all operands are non-negative here, so sum |a b| is the reference result itself
    assert_rows_within_bound(C.cpu().numpy(), C_ref, rph, C_ref.astype(np.float64), f"C3 k={k}")


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("k", [1, 5, 16, 32, 48, 64, 100, 128, 256])
@pytest.mark.parametrize("vt", [np.float32, np.float64])
def test_spmm_both_kernels_every_width(cuda, oracle, monkeypatch, variant, k, vt):
    """The group kernel (0) and the stream kernel (1) forced at every width: skewed row
    lengths (empty runs, a hub row spanning many streams), offsets with a non-zero base."""
    monkeypatch.setenv("SPBLAS_B200_SPMM_VARIANT", str(variant))
    rng = np.random.default_rng(1000 + k)
    m, n = 2311, 1777
    lens = rng.integers(0, 30, size=m)
    lens[rng.integers(0, m, size=m // 3)] = 0
    lens[100:400] = 0                                            # a long run of empty rows
    lens[1500] = 9000                                            # hub: shared by many streams
    lens[m - 1] = 77
    base = 123
    rp = (np.concatenate([[0], np.cumsum(lens)]) + base).astype(np.int32)
    nnz = int(rp[-1]) - base
    ci_full = rng.integers(0, n, size=nnz + base).astype(np.int32)
    v_full = rng.standard_normal(nnz + base).astype(vt)
    B = rng.standard_normal((n, k)).astype(vt)
    a = sb.csr_view(dev(v_full), dev(rp), dev(ci_full), (m, n), nnz)
    Bd = dev(B)
    Cd = torch.full((m, k), float("nan"), dtype=Bd.dtype, device="cuda")
    info = sb.multiply_inspect(a, Bd, Cd)
    sb.multiply_execute(info, sb.scaled(0.5, a), Bd, Cd)
    vec = 16 // np.dtype(vt).itemsize
    if variant == 1:                 # 16 bytes per lane once C is wider than 256 bytes
        assert info.spmm_variant == 1000 + (1 if (k * np.dtype(vt).itemsize <= 256 or k % vec) else vec)
    else:
        assert info.spmm_variant < 1000
    rp0 = (rp - base).astype(np.int32)
    ci, v = ci_full[base:], v_full[base:]
    C_ref = oracle.spmm("csr", (m, n), rp0, ci, v, B, alpha_a=0.5)
    assert_rows_within_bound(Cd.cpu().numpy(), C_ref, rp0, spmm_bound(rp0, ci, v, B, 0.5),
                             f"variant {variant} k={k}")
    # run to run: bit-identical (no atomics anywhere)
    C2 = torch.full_like(Cd, float("nan"))
    sb.multiply_execute(info, sb.scaled(0.5, a), Bd, C2)
    assert torch.equal(C2, Cd)
    info.close()


def test_row_length_histogram_steers_the_narrow_spmm(cuda, oracle):
    """A narrow B (rows of B shorter than 256 bytes) takes the group-per-row kernel on a matrix
    with even row lengths and the merge-path stream kernel when the inspect phase's row-length
    histogram says that >= 10 % of the entries sit in rows of >= 256 entries (a power-law
    matrix: 2-4x faster there, profiles/r02_spmm_rmat_narrow.jsonl).  Both within the bound."""
    rng = np.random.default_rng(77)
    m, n, k = 6000, 5000, 8
    for heavy in (False, True):
        lens = rng.integers(0, 24, size=m)
        if heavy:
            lens[rng.choice(m, 40, replace=False)] = rng.integers(300, 3000, 40)
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(rp[-1])
        ci = rng.integers(0, n, size=nnz).astype(np.int32)
        v = rng.standard_normal(nnz).astype(np.float32)
        B = rng.standard_normal((n, k)).astype(np.float32)
        a = csr_on_device(v, rp, ci, (m, n))
        Bd = dev(B)
        Cd = torch.full((m, k), float("nan"), dtype=Bd.dtype, device="cuda")
        info = sb.multiply_inspect(a, Bd, Cd)
        sb.multiply_execute(info, a, Bd, Cd)
        torch.cuda.synchronize()
        assert (info.spmm_variant >= 1000) == heavy, (heavy, info.spmm_variant)
        C_ref = oracle.spmm("csr", (m, n), rp, ci, v, B)
        assert_rows_within_bound(Cd.cpu().numpy(), C_ref, rp, spmm_bound(rp, ci, v, B),
                                 f"histogram-steered SpMM heavy={heavy}")
        info.close()
