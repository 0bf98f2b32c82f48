"""GPU parity tests for SpMV: the CUDA path (through the host API -> C ABI) against the
oracle, the committed golden vectors of the real reference, and — at BASELINE sizes —
size-independent properties.  Re-hosts the bodies of the reference's
test/gtest/device/spmv_test.cpp:11-146 (thrust_CsrView SpMV / SpMV_Ascaled / SpMV_BScaled)
and adds what the reference lacks: fp64, int32 scalars, int64 offsets, empty
rows, duplicates, very long rows, inspect reuse, multiply_execute, structure queries."""
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from conftest import DIMS, GOLDEN, golden
from helpers import (assert_rows_within_bound, csc_on_device, csr_on_device, dev, gpu_spmv)

pytestmark = pytest.mark.gpu
ALPHAS = [-10, 1, 5]


@pytest.mark.parametrize("dims", DIMS)
def test_reference_device_spmv_tests(cuda, oracle, dims):
    """thrust_CsrView.SpMV / SpMV_Ascaled / SpMV_BScaled on the reference's own fixtures,
    judged by the reference's own EXPECT_EQ_ and against the real reference's output."""
    g = golden(*dims)
    m, n, _ = dims
    v, rp, ci = g["csr_values"], g["csr_ptr"], g["csr_ind"]
    a = csr_on_device(v, rp, ci, (m, n))
    x = np.ones(n, np.float32)
    y = gpu_spmv(a, x, m, np.float32)
    assert oracle.expect_eq_tolerance(g["csr_spmv"], y).all()
    assert_rows_within_bound(y, g["csr_spmv"], rp, oracle.abs_rowsum(rp, ci, v, x), "spmv")
    for alpha in ALPHAS:
        ya = gpu_spmv(a, x, m, np.float32, alpha_a=alpha)
        assert oracle.expect_eq_tolerance(g[f"csr_spmv_ascaled_{alpha}"], ya).all()
        yb = gpu_spmv(a, x, m, np.float32, alpha_x=alpha)
        assert oracle.expect_eq_tolerance(g[f"csr_spmv_bscaled_{alpha}"], yb).all()
        bound = oracle.abs_rowsum(rp, ci, v, x, alpha)
        assert_rows_within_bound(ya, g[f"csr_spmv_ascaled_{alpha}"], rp, bound, "ascaled")
        assert_rows_within_bound(yb, g[f"csr_spmv_bscaled_{alpha}"], rp, bound, "bscaled")


def test_probe_semantics_on_gpu(cuda):
    p = np.load(f"{GOLDEN}/probe_3x4.npz")
    rp, ci, v, x = p["rowptr"], p["colind"], p["values"], p["x"]
    a = csr_on_device(v, rp, ci, (3, 4))
    assert gpu_spmv(a, x, 3, np.float32).tolist() == [140.0, 0.0, 160.0]   # stale NaN discarded
    xinf = x.copy()
    xinf[1] = np.inf                                                      # never referenced
    assert gpu_spmv(a, xinf, 3, np.float32).tolist() == [140.0, 0.0, 160.0]
    assert gpu_spmv(a, x, 3, np.float32, alpha_a=2, alpha_x=3).tolist() == [840.0, 0.0, 960.0]
    ai = csr_on_device(v.astype(np.int32), rp, ci, (3, 4))
    assert gpu_spmv(ai, x.astype(np.int32), 3, np.int32).tolist() == [140, 0, 160]
    # transposed(csr) behaves as the csc over the same arrays
    t = sb.transposed(a)
    assert gpu_spmv(t, np.ones(3, np.float32), 4, np.float32).tolist() == [2.0, 0.0, 4.0, 4.0]


def _random_csr(rng, m, n, kind, vt, it, ot):
    if kind == "short":
        lens = rng.integers(0, 12, size=m)
    elif kind == "mixed":
        lens = rng.integers(0, 6, size=m)
        lens[rng.integers(0, m, size=max(1, m // 50))] = rng.integers(65, 700, size=max(1, m // 50))
    elif kind == "hub":
        lens = rng.integers(0, 4, size=m)
        lens[m // 3] = 9000          # spans > 4 tiles of 2048 items
        lens[m - 1] = 2500
        lens[0] = 2049
    elif kind == "long":
        lens = rng.integers(100, 400, size=m)
    elif kind == "empty":
        lens = np.zeros(m, dtype=np.int64)
    else:
        raise ValueError(kind)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(ot)
    nnz = int(rp[-1])
    ci = rng.integers(0, n, size=nnz).astype(it)       # unsorted, duplicates legal
    if vt == np.int32:
        v = rng.integers(-9, 10, size=nnz).astype(vt)
        x = rng.integers(-9, 10, size=n).astype(vt)
    else:
        v = rng.standard_normal(nnz).astype(vt)
        x = rng.standard_normal(n).astype(vt)
    return v, rp, ci, x


@pytest.mark.parametrize("kind", ["short", "mixed", "hub", "long", "empty"])
@pytest.mark.parametrize("types", [(np.float32, np.int32, np.int32),
                                   (np.float64, np.int32, np.int32),
                                   (np.int32, np.int32, np.int32),
                                   (np.float64, np.int32, np.int64),
                                   (np.float32, np.int64, np.int64)])
def test_spmv_vs_oracle(cuda, oracle, kind, types):
    vt, it, ot = types
    rng = np.random.default_rng(zlib.crc32(f"{kind}{vt.__name__}{ot.__name__}".encode()))
    m, n = 3001, 1777
    v, rp, ci, x = _random_csr(rng, m, n, kind, vt, it, ot)
    a = csr_on_device(v, rp, ci, (m, n))
    alpha = 3 if vt == np.int32 else 0.75
    for kw in ({}, {"alpha_a": alpha}):
        y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x, **kw)
        bound = None if vt == np.int32 else oracle.abs_rowsum(rp, ci, v, x, kw.get("alpha_a", 1.0))
        # no-info path (light inspect per call), inspected path, multiply_execute spelling
        info = sb.multiply_inspect(a, dev(x), torch.empty(m, dtype=dev(x).dtype, device="cuda"))
        for y in (gpu_spmv(a, x, m, vt, **kw), gpu_spmv(a, x, m, vt, info=info, **kw),
                  gpu_spmv(a, x, m, vt, info=info, execute=True, **kw)):
            assert_rows_within_bound(y, y_ref, rp, bound, f"{kind} {vt.__name__}")
        info.close()


@pytest.mark.parametrize("kind", ["short", "hub", "empty"])
def test_inspect_structures_bit_exact(cuda, oracle, kind):
    """Row-length histogram, merge-path partition table and SpMM segments produced by the
    GPU inspect equal the CPU restatement exactly."""
    rng = np.random.default_rng(11)
    m, n = 5000, 4000
    v, rp, ci, x = _random_csr(rng, m, n, kind, np.float32, np.int32, np.int32)
    a = csr_on_device(v, rp, ci, (m, n))
    info = sb.multiply_inspect(a, dev(x), torch.empty(m, device="cuda"))
    hist, mx = oracle.rowlen_hist(rp)
    assert np.array_equal(info.rowlen_hist, hist)
    assert info.max_row_len == mx and info.empty_rows == hist[0]
    tile = info.tile_items
    want = oracle.merge_partition(rp, tile)
    assert info.num_tiles == len(want) - 1
    assert np.array_equal(info.tile_starts, want)
    assert np.array_equal(info.tile_uniform, oracle.tile_uniform(rp, want))
    segs = oracle.row_segments(rp, 4096)
    assert info.num_segments == len(segs)
    if len(segs):
        assert np.array_equal(info.segments, segs)
    assert info.last_launches == 0
    y = torch.empty(m, device="cuda")
    sb.multiply(info, a, dev(x), y)
    assert info.last_launches >= 1 and info.total_launches == info.last_launches
    info.close()


def test_values_may_change_between_executes(cuda, oracle):
    rng = np.random.default_rng(2)
    m, n = 2000, 2000
    v, rp, ci, x = _random_csr(rng, m, n, "short", np.float64, np.int32, np.int32)
    a = csr_on_device(v, rp, ci, (m, n))
    xd, y = dev(x), torch.empty(m, dtype=torch.float64, device="cuda")
    info = sb.multiply_inspect(a, xd, y)
    for rep in range(3):
        v2 = rng.standard_normal(len(v))
        a.values.copy_(dev(v2))
        sb.multiply_execute(info, a, xd, y)
        assert_rows_within_bound(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v2, x), rp,
                                 oracle.abs_rowsum(rp, ci, v2, x), "values changed")


def test_shard_with_nonzero_base_and_unaligned_pointers(cuda, oracle):
    """A row block of a larger matrix keeps the global rowptr values (base != 0) and slices
    of the global arrays whose addresses are not 16-byte aligned."""
    rng = np.random.default_rng(4)
    m, n = 4000, 3000
    v, rp, ci, x = _random_csr(rng, m, n, "mixed", np.float32, np.int32, np.int32)
    y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x)
    vd, rpd, cid, xd = dev(v), dev(rp), dev(ci), dev(x)
    r0, r1 = 1001, 3333
    # views index values/colind with the absolute offsets, exactly like
    # backend/view_customizations.hpp:48-66 does
    a = sb.csr_view(vd, rpd[r0:r1 + 1], cid, (r1 - r0, n), int(rp[r1] - rp[r0]))
    y = torch.full((r1 - r0,), float("nan"), device="cuda")
    info = sb.multiply_inspect(a, xd, y)
    sb.multiply(info, a, xd, y)
    bound = oracle.abs_rowsum(rp, ci, v, x)[r0:r1]
    assert_rows_within_bound(y.cpu().numpy(), y_ref[r0:r1], rp[r0:r1 + 1], bound, "shard")
    # unaligned: shift every array by one element
    pad = lambda t: torch.cat([t[:1], t])[1:]
    a2 = sb.csr_view(pad(vd), rpd, pad(cid), (m, n), int(rp[-1]))
    assert a2.values.data_ptr() % 16 != 0
    y2 = torch.empty(m, device="cuda")
    sb.multiply(a2, xd, y2)
    assert_rows_within_bound(y2.cpu().numpy(), y_ref, rp, oracle.abs_rowsum(rp, ci, v, x), "unaligned")


def test_errors_match_reference(cuda):
    p = np.load(f"{GOLDEN}/probe_3x4.npz")
    a = csr_on_device(p["values"], p["rowptr"], p["colind"], (3, 4))
    with pytest.raises(ValueError, match="matrix and vector dimensions are incompatible"):
        sb.multiply(a, torch.zeros(5, device="cuda"), torch.zeros(3, device="cuda"))
    with pytest.raises(RuntimeError):
        sb.multiply(a, torch.zeros(4, device="cuda", dtype=torch.float64),
                    torch.zeros(3, device="cuda"))
    bad = sb.csr_view(a.values, dev(np.array([0, 3, 2, 4], np.int32)), a.colind, (3, 4), 4)
    with pytest.raises(RuntimeError, match="monoton"):
        sb.multiply_inspect(bad, torch.zeros(4, device="cuda"), torch.zeros(3, device="cuda"))


def test_zero_sized(cuda):
    a = sb.csr_view(torch.zeros(0, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda"),
                    torch.zeros(0, dtype=torch.int32, device="cuda"), (0, 5), 0)
    sb.multiply(a, torch.zeros(5, device="cuda"), torch.zeros(0, device="cuda"))
    a = sb.csr_view(torch.zeros(0, device="cuda"), torch.zeros(8, dtype=torch.int32, device="cuda"),
                    torch.zeros(0, dtype=torch.int32, device="cuda"), (7, 5), 0)
    y = torch.full((7,), float("nan"), device="cuda")
    sb.multiply(a, torch.zeros(5, device="cuda"), y)
    assert y.cpu().tolist() == [0.0] * 7


# ---- BASELINE-size cases: generated on the device, checked against the oracle run on the
# host copy (the C oracle does 80M nonzeros in well under a second) and by properties ------
def test_c1_uniform_random_full_size(cuda, oracle):
    m = n = 1_000_000
    v, rp, ci, shape = G.uniform_random_csr(m, n, 10, seed=0, dtype=torch.float32, device=cuda)
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = torch.ones(n, device=cuda)
    y = torch.empty(m, device=cuda)
    info = sb.multiply_inspect(a, x, y)
    sb.multiply(info, sb.scaled(1.2, a), x, y)                  # examples/simple_spmv.cpp:44-48
    vh, rph, cih, xh = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy(), x.cpu().numpy()
    y_ref = oracle.spmv("csr", shape, rph, cih, vh, xh, alpha_a=1.2)
    assert_rows_within_bound(y.cpu().numpy(), y_ref, rph, oracle.abs_rowsum(rph, cih, vh, xh, 1.2), "C1")


def test_c2_poisson_known_answer_and_iteration(cuda, oracle):
    g = 1024                                                   # full 4096 runs in bench.py
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, cuda)
    n = g * g
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = torch.ones(n, dtype=torch.float64, device=cuda)
    y = torch.empty(n, dtype=torch.float64, device=cuda)
    info = sb.multiply_inspect(a, x, y)
    tu = info.tile_uniform
    assert np.array_equal(tu, oracle.tile_uniform(rp.cpu().numpy(), info.tile_starts))
    assert (tu == 5).mean() > 0.6                          # interior tiles take the stencil path
    sb.multiply(info, a, x, y)
    # A * 1 = 4 - (number of neighbours): exact small integers
    i, j = torch.arange(n, device=cuda) // g, torch.arange(n, device=cuda) % g
    want = 4.0 - ((i > 0).double() + (i < g - 1).double() + (j > 0).double() + (j < g - 1).double())
    assert torch.equal(y, want)
    # iterated y -> x with scaled(1/8, a), checked per iteration against the oracle fed the
    # same input (not compounded), SURVEY §8d
    x = G.dense_uniform((n,), 1, torch.float64, cuda)
    vh, rph, cih = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy()
    for it in range(3):
        sb.multiply_execute(info, sb.scaled(0.125, a), x, y)
        xh = x.cpu().numpy()
        y_ref = oracle.spmv("csr", shape, rph, cih, vh, xh, alpha_a=0.125)
        assert_rows_within_bound(y.cpu().numpy(), y_ref, rph,
                                 oracle.abs_rowsum(rph, cih, vh, xh, 0.125), f"C2 it{it}")
        x, y = y, x


def test_c4_rmat_skewed_rows(cuda, oracle):
    scale = 18                                                 # scale 24 runs in bench.py
    v, rp, ci, shape = G.rmat_csr(scale, 16, seed=24, dtype=torch.float32, device=cuda)
    m, n = shape
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = G.dense_uniform((n,), 5, torch.float32, cuda)
    y = torch.empty(m, device=cuda)
    info = sb.multiply_inspect(a, x, y)
    assert info.max_row_len > 2048                             # hubs span several tiles
    sb.multiply(info, a, x, y)
    vh, rph, cih, xh = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy(), x.cpu().numpy()
    y_ref = oracle.spmv("csr", shape, rph, cih, vh, xh)
    worst = assert_rows_within_bound(y.cpu().numpy(), y_ref, rph,
                                     oracle.abs_rowsum(rph, cih, vh, xh), "C4")
    assert worst <= 1.0
    hist, mx = oracle.rowlen_hist(rph)
    assert np.array_equal(info.rowlen_hist, hist) and info.max_row_len == mx
    assert np.array_equal(info.tile_starts, oracle.merge_partition(rph, info.tile_items))
    # linearity: A(2x + x) == 3 A x up to the same bound
    y3 = torch.empty(m, device=cuda)
    sb.multiply(info, a, 3 * x, y3)
    assert_rows_within_bound(y3.cpu().numpy(), 3 * y_ref, rph,
                             3 * oracle.abs_rowsum(rph, cih, vh, xh), "C4 linear")


def test_csc_spmv(cuda, oracle):
    """csc_view SpMV (column-major storage): golden fixtures and a random case; the
    inspect phase builds the row-major image bit-exactly as the oracle does."""
    for dims in DIMS:
        g = golden(*dims)
        m, n, _ = dims
        v, cp, ri = g["csc_values"], g["csc_ptr"], g["csc_ind"]
        a = csc_on_device(v, cp, ri, (m, n))
        x = np.ones(n, np.float32)
        y = gpu_spmv(a, x, m, np.float32)
        assert oracle.expect_eq_tolerance(g["csc_spmv"], y).all()
        xd, yd = dev(x), torch.empty(m, device="cuda")
        info = sb.multiply_inspect(a, xd, yd)
        t_rp, t_ci, perm = oracle.csc_row_major_image((m, n), cp, ri)
        g_rp, g_ci, g_perm = info.effective_csr(np.int32, np.int32)
        assert np.array_equal(g_rp, t_rp) and np.array_equal(g_ci, t_ci)
        assert np.array_equal(g_perm, perm)
        # same addition order as the reference's column scatter -> compare tightly
        sb.multiply(info, sb.scaled(5, a), xd, yd)
        assert oracle.expect_eq_tolerance(g["csc_spmv_ascaled_5"], yd.cpu().numpy()).all()


@pytest.mark.parametrize("ws_items", ["256", "0"])
@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("kind", ["short", "mixed", "hub", "long", "empty"])
def test_spmv_every_kernel_variant(cuda, oracle, monkeypatch, kind, variant, ws_items):
    """Each SpMV kernel (0 one tile per CTA, 1 TMA pipeline, 2 warp streams) forced on every
    row-length mix, in fp32 / fp64 / int32 (exact), on a CSR matrix, on a row-block shard
    with a non-zero base and on a CSC matrix (value permutation).  ws_items = 256 makes a
    warp walk many short streams (carries between streams, rows spanning several)."""
    if variant != 2 and ws_items == "0":
        pytest.skip("stream length only matters to the warp-stream kernel")
    monkeypatch.setenv("SPBLAS_B200_SPMV_VARIANT", str(variant))
    if ws_items != "0":
        monkeypatch.setenv("SPBLAS_B200_WS_ITEMS", ws_items)
    for vt in (np.float32, np.float64, np.int32):
        rng = np.random.default_rng(zlib.crc32(f"var{kind}{vt.__name__}".encode()))
        m, n = 5003, 2777
        v, rp, ci, x = _random_csr(rng, m, n, kind, vt, np.int32, np.int32)
        alpha = 3 if vt == np.int32 else 0.75
        y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=alpha)
        bound = None if vt == np.int32 else oracle.abs_rowsum(rp, ci, v, x, alpha)
        a = csr_on_device(v, rp, ci, (m, n))
        xd = dev(x)
        info = sb.multiply_inspect(a, xd, torch.empty(m, dtype=xd.dtype, device="cuda"))
        y = gpu_spmv(a, x, m, vt, info=info, alpha_a=alpha)
        assert info.spmv_variant == variant
        assert_rows_within_bound(y, y_ref, rp, bound, f"variant {variant} {kind} {vt.__name__}")
        info.close()
        # a row block of the same matrix: rowptr keeps the global base
        r0, r1 = 777, 4100
        blk = sb.csr_view(a.values, a.rowptr[r0:r1 + 1], a.colind, (r1 - r0, n),
                          int(rp[r1] - rp[r0]))
        info = sb.multiply_inspect(blk, xd, torch.empty(r1 - r0, dtype=xd.dtype, device="cuda"))
        yb = gpu_spmv(blk, x, r1 - r0, vt, info=info, alpha_a=alpha)
        assert_rows_within_bound(yb, y_ref[r0:r1], rp[r0:r1 + 1],
                                 None if bound is None else bound[r0:r1], f"shard variant {variant}")
        info.close()
    # CSC: the kernels gather the values through the permutation built by the inspect
    rng = np.random.default_rng(zlib.crc32(f"varcsc{kind}".encode()))
    m, n = 2203, 3001
    v, cp, ri, _ = _random_csr(rng, n, m, kind, np.float64, np.int32, np.int32)
    x = rng.standard_normal(n)
    ac = csc_on_device(v, cp, ri, (m, n))
    info = sb.multiply_inspect(ac, dev(x), torch.empty(m, dtype=torch.float64, device="cuda"))
    yc = gpu_spmv(ac, x, m, np.float64, info=info)
    t_rp, t_ci, perm = oracle.csc_row_major_image((m, n), cp, ri)
    assert_rows_within_bound(yc, oracle.spmv("csc", (m, n), cp, ri, v, x), t_rp,
                             oracle.abs_rowsum(t_rp, t_ci, v[perm], x), f"csc variant {variant}")
    info.close()


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_matrix_opt_caches_values_of_a_transposed_operand(cuda, oracle, monkeypatch, variant):
    """multiply_inspect(matrix_opt(transposed(a))) lets the plan keep the values gathered in
    image order (spblas_b200_plan_cache_values): executes are bit-identical to the uncached
    path; as with the reference's oneMKL optimize, changed values are only seen after the
    next multiply_inspect; an operand without matrix_opt is never cached."""
    monkeypatch.setenv("SPBLAS_B200_SPMV_VARIANT", str(variant))
    rng = np.random.default_rng(31 + variant)
    m, n = 2777, 3501
    v, rp, ci, _ = _random_csr(rng, m, n, "mixed", np.float64, np.int32, np.int32)
    x = rng.standard_normal(m)
    a = csr_on_device(v, rp, ci, (m, n))
    at = sb.transposed(a)                                       # n x m, CSC over A's arrays
    xd = dev(x)
    y_plain = torch.empty(n, dtype=torch.float64, device="cuda")
    info_plain = sb.multiply_inspect(at, xd, y_plain)
    sb.multiply_execute(info_plain, sb.scaled(0.5, at), xd, y_plain)
    y_opt = torch.full((n,), float("nan"), dtype=torch.float64, device="cuda")
    aopt = sb.matrix_opt(at)
    info = sb.multiply_inspect(aopt, xd, y_opt)
    sb.multiply_execute(info, sb.scaled(0.5, aopt), xd, y_opt)
    torch.cuda.synchronize()
    assert torch.equal(y_plain, y_opt)
    t_rp, t_ci, perm = oracle.csc_row_major_image((n, m), rp, ci)
    y_ref = oracle.spmv("csc", (n, m), rp, ci, v, x, alpha_a=0.5)
    assert_rows_within_bound(y_opt.cpu().numpy(), y_ref, t_rp,
                             oracle.abs_rowsum(t_rp, t_ci, v[perm], x, 0.5), "matrix_opt transposed")
    # SpMM through the same cached plan state
    B = rng.standard_normal((m, 8))
    C1, C2 = (torch.empty((n, 8), dtype=torch.float64, device="cuda") for _ in range(2))
    i1, i2 = sb.multiply_inspect(at, dev(B), C1), sb.multiply_inspect(aopt, dev(B), C2)
    sb.multiply(i1, at, dev(B), C1)
    sb.multiply(i2, aopt, dev(B), C2)
    assert torch.equal(C1, C2)
    # new values: the plain plan follows at once, the optimised one after re-inspection
    v2 = rng.standard_normal(len(v))
    a.values.copy_(dev(v2))
    sb.multiply_execute(info_plain, at, xd, y_plain)
    sb.multiply_inspect(info, aopt, xd, y_opt)
    sb.multiply_execute(info, aopt, xd, y_opt)
    torch.cuda.synchronize()
    assert torch.equal(y_plain, y_opt)
    assert_rows_within_bound(y_opt.cpu().numpy(), oracle.spmv("csc", (n, m), rp, ci, v2, x), t_rp,
                             oracle.abs_rowsum(t_rp, t_ci, v2[perm], x), "matrix_opt re-inspected")
    for i in (info_plain, info, i1, i2):
        i.close()


@pytest.mark.parametrize("stages", ["2", "3", "4", "6", "8"])
def test_pipe_kernel_ragged_last_tile_every_stage_count(cuda, oracle, monkeypatch, stages):
    """The pipelined kernel's last tile mixes plain stores (zero fill past the arrays' end),
    element-wise cp.async (the last partial quad, the row ends) and TMA bulk copies on ONE
    stage of the ring; racecheck cannot model their completion mechanism (cp.async.mbarrier.
    arrive.noinc), so this settles it by execution: nnz % 4 = 0..3, rows uniform (exact path)
    and mixed (flat path), every ring depth, 25 executes each into a poisoned y — integer
    scalars, so a single stale or torn operand shows as an exact mismatch."""
    monkeypatch.setenv("SPBLAS_B200_SPMV_VARIANT", "1")
    monkeypatch.setenv("SPBLAS_B200_STAGES", stages)
    rng = np.random.default_rng(int(stages))
    n = 1531
    for uniform in (True, False):
        for tail in range(4):
            # a few tiles' worth of rows (2048 merge items per tile), then trim to nnz % 4 == tail
            m = 3 * 2048 // 6 + 17
            lens = np.full(m, 5, dtype=np.int64) if uniform else rng.integers(0, 11, size=m)
            while int(lens.sum()) % 4 != tail:
                lens[-1 - int(rng.integers(0, 3))] += 1
            rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            nnz = int(rp[-1])
            ci = rng.integers(0, n, size=nnz).astype(np.int32)
            v = rng.integers(-9, 10, size=nnz).astype(np.int32)
            x = rng.integers(-9, 10, size=n).astype(np.int32)
            want = oracle.spmv("csr", (m, n), rp, ci, v, x)
            # values / colind end exactly at the allocation's end: a read past nnz would fault
            a = csr_on_device(v, rp, ci, (m, n))
            xd = dev(x)
            y = torch.empty(m, dtype=torch.int32, device="cuda")
            info = sb.multiply_inspect(a, xd, y)
            for rep in range(25):
                y.fill_(77)
                sb.multiply_execute(info, a, xd, y)
                assert info.spmv_variant == 1
                assert np.array_equal(y.cpu().numpy(), want), (uniform, tail, stages, rep)
            info.close()


@pytest.mark.parametrize("ws_items", ["256", "1024", "0"])
@pytest.mark.parametrize("variant", [2, 3, 4])
def test_walk_lanes_outside_their_stream_contribute_nothing(cuda, oracle, monkeypatch, variant,
                                                            ws_items):
    """The walk's branch-free load phase lets a lane whose quad lies outside [k, kend) of its
    stream load the chunk's first quad instead (entries of ANOTHER stream's rows) and zeroes
    its products.  x holds NaN at columns that only one marked row references: if a product
    taken outside a lane's own range leaked — or were multiplied by zero instead of being
    dropped — a second row would turn NaN.  Streams of 256 / 1024 entries start and end in the
    middle of chunks and of rows; variants 3 and 4 read the hub tables (re-encoded colind)."""
    monkeypatch.setenv("SPBLAS_B200_SPMV_VARIANT", str(variant))
    if ws_items != "0":
        monkeypatch.setenv("SPBLAS_B200_WS_ITEMS", ws_items)
    rng = np.random.default_rng(100 + variant)
    m, n = 6007, 5000
    n_clean = 4000                                   # columns 4000.. are referenced by marked rows only
    lens = rng.integers(0, 40, m)
    lens[rng.integers(0, m, 12)] = rng.integers(300, 900, 12)      # rows spanning several chunks
    marked = np.sort(rng.choice(m, 25, replace=False))
    rp = np.zeros(m + 1, np.int32)
    np.cumsum(lens, out=rp[1:])
    # skewed columns so that hub tables form (variants 3 / 4 keep them only if they pay)
    ci = np.minimum((n_clean * rng.random(rp[-1]) ** 3).astype(np.int32), n_clean - 1)
    v = rng.standard_normal(rp[-1])
    x = rng.standard_normal(n)
    for t, r in enumerate(marked):
        if lens[r] == 0:
            continue
        ci[rp[r] + rng.integers(0, lens[r])] = n_clean + t          # one private column each
        x[n_clean + t] = np.nan
    for vt in (np.float32, np.float64):
        a = csr_on_device(v.astype(vt), rp, ci, (m, n))
        xd = dev(x.astype(vt))
        info = sb.multiply_inspect(a, xd, torch.empty(m, dtype=xd.dtype, device="cuda"))
        if variant >= 3:
            info.set_hub(True, 2048, 2)
            info.force_spmv_variant(variant)
        y = gpu_spmv(a, x.astype(vt), m, vt, info=info)
        assert info.spmv_variant == variant
        expect_nan = np.zeros(m, bool)
        expect_nan[marked[lens[marked] > 0]] = True
        assert np.array_equal(np.isnan(y), expect_nan), \
            f"variant {variant}: NaN rows {np.flatnonzero(np.isnan(y) != expect_nan)[:8]}"
        xz = np.where(np.isnan(x), 0.0, x).astype(vt)
        y_ref = oracle.spmv("csr", (m, n), rp, ci, v.astype(vt), xz)
        ok = ~expect_nan
        bound = oracle.abs_rowsum(rp, ci, v.astype(vt), xz)
        err = np.abs(y.astype(np.float64) - y_ref)[ok]
        tol = ((lens + 2) * (2.0 ** -23 if vt == np.float32 else 2.0 ** -52) * bound)[ok]
        assert (err <= tol).all(), f"variant {variant} {vt.__name__}: max err/tol {np.max(err / np.maximum(tol, 1e-300))}"
        info.close()


def test_alternating_a_and_transposed_a_with_one_info(cuda, oracle, monkeypatch):
    """The reference's flagship iterative use (notes/spmv.hpp:12-22): ONE operation_info_t is
    inspected for `a` and for `transposed(a)`, then the executes alternate.  The info keeps a
    plan per structure: after the two inspects no execute inspects again, and every product is
    within the bound of the reference's CPU multiply on the same input."""
    import sys
    M = sys.modules["spblas_reference_b200.multiply"]
    rng = np.random.default_rng(2024)
    m, n = 3001, 2500
    v, rp, ci, _ = _random_csr(rng, m, n, "mixed", np.float64, np.int32, np.int32)
    v = v * 0.05
    a = csr_on_device(v, rp, ci, (m, n))
    at = sb.transposed(a)
    x = rng.standard_normal(n)
    xd, yd = dev(x), torch.empty(m, dtype=torch.float64, device="cuda")
    info = sb.operation_info_t()
    sb.multiply_inspect(info, a, xd, yd)
    sb.multiply_inspect(info, at, yd, xd)
    calls = []
    real = M._inspect
    monkeypatch.setattr(M, "_inspect", lambda *args, **kw: (calls.append(1), real(*args, **kw))[1])
    cp, ri = rp, ci                                  # transposed(a): the CSC of A^T over the same arrays
    t_rp, t_ci, perm = oracle.csc_row_major_image((n, m), cp, ri)
    for it in range(3):
        sb.multiply_execute(info, a, xd, yd)
        torch.cuda.synchronize()
        assert info.spmv_variant in (0, 1, 2)
        y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x)
        assert_rows_within_bound(yd.cpu().numpy(), y_ref, rp, oracle.abs_rowsum(rp, ci, v, x),
                                 f"A x, iteration {it}")
        y = yd.cpu().numpy()
        sb.multiply_execute(info, at, yd, xd)
        torch.cuda.synchronize()
        x_ref = oracle.spmv("csc", (n, m), cp, ri, v, y)
        assert_rows_within_bound(xd.cpu().numpy(), x_ref, t_rp,
                                 oracle.abs_rowsum(t_rp, t_ci, v[perm], y), f"A^T y, iteration {it}")
        x = xd.cpu().numpy()
    assert not calls, "an execute re-inspected although both structures had been inspected"
    info.close()
