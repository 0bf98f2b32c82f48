"""Pins the oracle (oracle/spblas_oracle.c): against the real reference compiled from
/root/reference (oracle/_ref, when present), against the committed golden vectors the real
reference produced, and with the reference tests' own known-answer loops and tolerance
(test/gtest/spmv_test.cpp:21-34, spmm_test.cpp:26-41, util.hpp:7-23).  CPU only."""
import numpy as np
import pytest

from conftest import DIMS, GOLDEN, golden

ALPHAS = [-10, 1, 5]


def _known_answer_spmv(m, rowptr, colind, values, x, alpha=1):
    # the triple loop of test/gtest/spmv_test.cpp:21-30
    ref = np.zeros(m, dtype=values.dtype)
    for i in range(m):
        for p in range(rowptr[i], rowptr[i + 1]):
            ref[i] += values.dtype.type(alpha) * values[p] * x[colind[p]]
    return ref


@pytest.mark.parametrize("dims", DIMS)
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_spmv_matches_golden_bit_exact(oracle, dims, fmt):
    g = golden(*dims)
    m, n, _ = dims
    v, ptr, ind = g[f"{fmt}_values"], g[f"{fmt}_ptr"], g[f"{fmt}_ind"]
    x = np.ones(n, np.float32)
    assert np.array_equal(oracle.spmv(fmt, (m, n), ptr, ind, v, x), g[f"{fmt}_spmv"])
    for a in ALPHAS:
        assert np.array_equal(oracle.spmv(fmt, (m, n), ptr, ind, v, x, alpha_a=a),
                              g[f"{fmt}_spmv_ascaled_{a}"])
        assert np.array_equal(oracle.spmv(fmt, (m, n), ptr, ind, v, x, alpha_x=a),
                              g[f"{fmt}_spmv_bscaled_{a}"])


@pytest.mark.parametrize("dims", DIMS)
@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("k", [1, 8, 32, 64, 512])
def test_spmm_matches_golden_bit_exact(oracle, dims, fmt, k):
    g = golden(*dims)
    m, n, _ = dims
    v, ptr, ind = g[f"{fmt}_values"], g[f"{fmt}_ptr"], g[f"{fmt}_ind"]
    B = g[f"dense_B_{k}"]
    assert np.array_equal(oracle.spmm(fmt, (m, n), ptr, ind, v, B), g[f"{fmt}_spmm_{k}"])
    if k <= 64:
        assert np.array_equal(oracle.spmm(fmt, (m, n), ptr, ind, v, B, alpha_a=2.0),
                              g[f"{fmt}_spmm_ascaled_{k}"])


@pytest.mark.parametrize("dims", DIMS)
def test_reference_known_answer_loop(oracle, dims):
    """The reference's own acceptance test re-hosted: triple loop + EXPECT_EQ_."""
    g = golden(*dims)
    m, n, _ = dims
    v, rp, ci = g["csr_values"], g["csr_ptr"], g["csr_ind"]
    x = np.ones(n, np.float32)
    for alpha in [1] + ALPHAS:
        y = oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=None if alpha == 1 else alpha)
        ka = _known_answer_spmv(m, rp, ci, v, x, alpha)
        assert oracle.expect_eq_tolerance(ka, y).all()


def test_probe_semantics(oracle):
    """SURVEY Appendix A probe: duplicates accumulate, empty rows give 0, stale NaN in y is
    discarded (oracle.spmv pre-fills y with NaN), unreferenced Inf does not propagate."""
    p = np.load(f"{GOLDEN}/probe_3x4.npz")
    rp, ci, v, x = p["rowptr"], p["colind"], p["values"], p["x"]
    assert oracle.spmv("csr", (3, 4), rp, ci, v, x).tolist() == [140.0, 0.0, 160.0]
    xinf = x.copy()
    xinf[1] = np.inf
    assert np.array_equal(oracle.spmv("csr", (3, 4), rp, ci, v, xinf), p["y_inf"])
    assert np.array_equal(oracle.spmv("csr", (3, 4), rp, ci, v, x, alpha_a=2, alpha_x=3),
                          p["y_scaled"])
    assert np.array_equal(oracle.spmv("csc", (4, 3), rp, ci, v, np.ones(3, np.float32)),
                          p["yt"])
    assert np.array_equal(oracle.spmv("csr", (3, 4), rp, ci, v.astype(np.int32),
                                      x.astype(np.int32)), p["y_s32"])


def test_against_real_reference_when_present(oracle):
    """Bit-exact against the real spblas::multiply for every type combination the shim
    instantiates, including the inspect + multiply(info, ...) spelling."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(7)
    for (vt, it, ot) in [(np.float32, np.int32, np.int32), (np.float32, np.int32, np.int64),
                         (np.float64, np.int32, np.int32), (np.float64, np.int32, np.int64),
                         (np.int32, np.int32, np.int32), (np.float32, np.int64, np.int64)]:
        m, n = 257, 131
        lens = rng.integers(0, 40, size=m)
        lens[5] = 0
        lens[17] = 300
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(ot)
        ci = rng.integers(0, n, size=int(rp[-1])).astype(it)
        if vt == np.int32:
            v = rng.integers(-9, 9, size=len(ci)).astype(vt)
            x = rng.integers(-9, 9, size=n).astype(vt)
            B = rng.integers(-9, 9, size=(n, 9)).astype(vt)
            aa, ax = 3, -2
        else:
            v = rng.standard_normal(len(ci)).astype(vt)
            x = rng.standard_normal(n).astype(vt)
            B = rng.standard_normal((n, 9)).astype(vt)
            aa, ax = 1.5, -0.25
        for insp in (False, True):
            for kw in ({}, {"alpha_a": aa}, {"alpha_x": ax}, {"alpha_a": aa, "alpha_x": ax}):
                a = oracle.spmv("csr", (m, n), rp, ci, v, x, **kw)
                b = oracle.spmv("csr", (m, n), rp, ci, v, x, impl="reference", inspect=insp, **kw)
                assert np.array_equal(a, b), (vt, it, ot, kw)
        # CSC: the same arrays read as the transpose (n x m matrix... here m x n with
        # colptr over m "columns"): shape (n, m)
        xt = (rng.standard_normal(m).astype(vt) if vt != np.int32
              else rng.integers(-9, 9, size=m).astype(vt))
        a = oracle.spmv("csc", (n, m), rp, ci, v, xt)
        b = oracle.spmv("csc", (n, m), rp, ci, v, xt, impl="reference")
        assert np.array_equal(a, b)
        for kw in ({}, {"alpha_a": aa}, {"alpha_b": ax}):
            a = oracle.spmm("csr", (m, n), rp, ci, v, B, **kw)
            b = oracle.spmm("csr", (m, n), rp, ci, v, B, impl="reference", **kw)
            assert np.array_equal(a, b), (vt, kw)
        Bt = (rng.standard_normal((m, 5)).astype(vt) if vt != np.int32
              else rng.integers(-9, 9, size=(m, 5)).astype(vt))
        a = oracle.spmm("csc", (n, m), rp, ci, v, Bt)
        b = oracle.spmm("csc", (n, m), rp, ci, v, Bt, impl="reference")
        assert np.array_equal(a, b)


def test_reference_fixtures_match_golden(oracle):
    """generate_csr through the shim reproduces the committed fixture arrays."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for (m, n, nnz) in DIMS:
        g = golden(m, n, nnz)
        v, rp, ci = oracle.ref_generate_csr(m, n, nnz)
        assert np.array_equal(v, g["csr_values"])
        assert np.array_equal(rp, g["csr_ptr"])
        assert np.array_equal(ci, g["csr_ind"])


# ---- inspect-phase restatements -------------------------------------------------------
def _random_rowptr(rng, rows, kind):
    if kind == "uniform":
        lens = rng.integers(0, 12, size=rows)
    elif kind == "empty":
        lens = np.zeros(rows, dtype=np.int64)
    elif kind == "hub":
        lens = rng.integers(0, 4, size=rows)
        lens[rows // 3] = 9000
        lens[rows - 1] = 2500
    else:
        lens = np.full(rows, 5)
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


@pytest.mark.parametrize("kind", ["uniform", "empty", "hub", "const"])
@pytest.mark.parametrize("tile", [7, 64, 2048])
def test_merge_partition_properties(oracle, kind, tile):
    rng = np.random.default_rng(3)
    rp = _random_rowptr(rng, 1500, kind) + 11          # non-zero base, as a shard has
    rows, nnz = len(rp) - 1, int(rp[-1] - rp[0])
    st = oracle.merge_partition(rp, tile)
    assert st[0].tolist() == [0, 11] and st[-1].tolist() == [rows, 11 + nnz]
    d = (st[:, 0] + st[:, 1] - 11)
    want = np.minimum(np.arange(len(st)) * tile, rows + nnz)
    assert np.array_equal(d, want)                      # every tile starts on its diagonal
    assert (np.diff(st[:, 0]) >= 0).all() and (np.diff(st[:, 1]) >= 0).all()
    # merge order: rows before the split are finished, the split row is not over-consumed
    for r, k in st:
        if r > 0:
            assert rp[r] <= k
        if r < rows:
            assert k <= rp[r + 1]


def test_tile_uniform(oracle):
    rp = np.arange(0, 5 * 1001, 5, dtype=np.int64)           # 1000 rows of 5
    st = oracle.merge_partition(rp, 64)
    tu = oracle.tile_uniform(rp, st)
    assert (tu == 5).all()
    rp2 = rp.copy()
    rp2[500:] += 1                                            # row 499 has 6 entries
    st2 = oracle.merge_partition(rp2, 64)
    tu2 = oracle.tile_uniform(rp2, st2)
    bad = [t for t in range(len(tu2)) if st2[t, 0] + 1 <= 499 < st2[t + 1, 0]]
    assert len(bad) == 1 and tu2[bad[0]] == 0 and (np.delete(tu2, bad) == 5).all()
    rp3 = np.arange(0, 9 * 200, 9, dtype=np.int64)            # rows of 9: longer than the cap
    assert (oracle.tile_uniform(rp3, oracle.merge_partition(rp3, 64)) == 0).all()


def test_rowlen_hist_and_segments(oracle):
    rng = np.random.default_rng(5)
    rp = _random_rowptr(rng, 4000, "hub")
    hist, mx = oracle.rowlen_hist(rp)
    lens = np.diff(rp)
    assert hist.sum() == len(lens) and mx == lens.max()
    assert hist[0] == (lens == 0).sum() and hist[1] == (lens == 1).sum()
    assert hist[2] == ((lens >= 2) & (lens < 4)).sum()
    assert hist[14] == ((lens >= 8192) & (lens < 16384)).sum() == 1
    segs = oracle.row_segments(rp, 4096)
    assert segs[:, 0].tolist() == [4000 // 3] * 3
    assert segs[0, 1] == rp[4000 // 3] and segs[-1, 2] == rp[4000 // 3 + 1]
    assert (segs[:, 2] - segs[:, 1]).tolist() == [4096, 4096, 9000 - 8192]
    assert oracle.rowlen_hist(np.array([0, 5, 3]))[1] == -1     # not monotone


def test_csc_row_major_image(oracle):
    g = golden(100, 1000, 10000)
    cp, ri, v = g["csc_ptr"], g["csc_ind"], g["csc_values"]
    t_rp, t_ci, perm = oracle.csc_row_major_image((100, 1000), cp, ri)
    assert t_rp[-1] == len(ri)
    # the image is the same matrix: CSR product on the image == CSC product
    x = np.arange(1000, dtype=np.float32) % 7
    y_csc = oracle.spmv("csc", (100, 1000), cp, ri, v, x)
    y_img = oracle.spmv("csr", (100, 1000), t_rp.astype(np.int32), t_ci.astype(np.int32),
                        v[perm], x)
    assert np.array_equal(y_csc, y_img)                 # same order of additions: bit-exact
    for i in range(100):
        seg = perm[t_rp[i]:t_rp[i + 1]]
        assert (np.diff(seg) > 0).all()                 # stable: storage order kept


# ---------------------------------------------------------------------------------------
# transpose(a, b): oracle restatement of algorithms/transpose_impl.hpp:14-53
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", DIMS)
def test_transpose_matches_golden_and_reference_test(oracle, dims):
    """Bit-exact against the committed output of the real spblas::transpose, and the
    reference test's own check (test/gtest/transpose_test.cpp:36-87: the COO triples of B,
    rows and columns swapped, are a permutation of A's)."""
    g = golden(*dims)
    m, n, _ = dims
    v, rp, ci = g["csr_values"], g["csr_ptr"], g["csr_ind"]
    tv, trp, tci = oracle.transpose((m, n), rp, ci, v)
    assert np.array_equal(tv, g["csr_transpose_values"])
    assert np.array_equal(trp, g["csr_transpose_ptr"])
    assert np.array_equal(tci, g["csr_transpose_ind"])
    a_rows = np.repeat(np.arange(m), np.diff(rp))
    b_rows = np.repeat(np.arange(n), np.diff(trp))
    a_coo = sorted(zip(ci.tolist(), a_rows.tolist(), v.tolist()))
    b_coo = sorted(zip(b_rows.tolist(), tci.tolist(), tv.tolist()))
    assert a_coo == b_coo
    assert trp[0] == 0 and trp[-1] == len(ci)


def test_transpose_against_real_reference_when_present(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(12)
    for (vt, it, ot) in [(np.float32, np.int32, np.int32), (np.float32, np.int32, np.int64),
                         (np.float64, np.int32, np.int32), (np.float64, np.int32, np.int64),
                         (np.int32, np.int32, np.int32), (np.float32, np.int64, np.int64)]:
        m, n = 311, 97
        lens = rng.integers(0, 30, size=m)
        lens[3] = 0
        lens[200] = 250                                  # duplicates inside a row are certain
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(ot)
        ci = rng.integers(0, n, size=int(rp[-1])).astype(it)          # unsorted
        v = (rng.integers(-99, 99, size=len(ci)) if vt == np.int32
             else rng.standard_normal(len(ci))).astype(vt)
        got = oracle.transpose((m, n), rp, ci, v)
        want = oracle.transpose((m, n), rp, ci, v, impl="reference")
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    # empty matrix, and a matrix with no entries in some columns
    e = oracle.transpose((4, 3), np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32))
    assert e[1].tolist() == [0, 0, 0, 0]


def test_transpose_equals_csc_image(oracle):
    """transpose(A) is the row-major image of A^T given as CSC over A's arrays — the identity
    the GPU implementation rests on (csrc/inspect.cu: run_transpose)."""
    rng = np.random.default_rng(13)
    m, n = 150, 220
    lens = rng.integers(0, 9, size=m)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = rng.integers(0, n, size=int(rp[-1])).astype(np.int32)
    v = rng.standard_normal(len(ci)).astype(np.float32)
    tv, trp, tci = oracle.transpose((m, n), rp, ci, v)
    i_rp, i_ci, perm = oracle.csc_row_major_image((n, m), rp, ci)
    assert np.array_equal(trp, i_rp) and np.array_equal(tci, i_ci) and np.array_equal(tv, v[perm])


# ---------------------------------------------------------------------------------------
# triangular_solve: oracle restatement of triangular_solve_impl.hpp:44-94
# ---------------------------------------------------------------------------------------
SQUARE_DIMS = [(1000, 1000, 100), (100, 100, 100), (40, 40, 1000)]        # util.hpp:31-33


def _reference_test_trsv(m, rowptr, colind, values, b, upper, unit):
    # reference_triangular_solve of test/gtest/triangular_solve_test.cpp:6-58 (tmp = b - sum)
    x = np.zeros(m, dtype=values.dtype)
    for row in (range(m - 1, -1, -1) if upper else range(m)):
        tmp, diag = values.dtype.type(b[row]), values.dtype.type(0)
        for j in range(rowptr[row], rowptr[row + 1]):
            col = colind[j]
            if (col > row) if upper else (col < row):
                tmp = values.dtype.type(tmp - values.dtype.type(values[j] * x[col]))
            elif col == row:
                diag = values[j]
        x[row] = tmp if unit else values.dtype.type(tmp / diag)
    return x


@pytest.mark.parametrize("dims", SQUARE_DIMS)
def test_trsv_matches_golden_and_reference_test(oracle, dims):
    g = np.load(f"{GOLDEN}/trsv_square_dims.npz")
    m, _, nnz = dims
    key = f"{m}_{nnz}"
    v, rp, ci, b = g[f"values_{key}"], g[f"ptr_{key}"], g[f"ind_{key}"], g[f"b_{key}"]
    dv, drp, dci = g[f"dvalues_{key}"], g[f"dptr_{key}"], g[f"dind_{key}"]
    for upper, name in ((0, "lower"), (1, "upper")):
        x = oracle.trsv(m, rp, ci, v, b, upper=upper, unit=True)
        assert np.array_equal(x, g[f"x_unit_{name}_{key}"])
        # the reference test's own solver (b - sum formed the other way round: a different
        # rounding sequence, amplified along the substitution) agrees to a few 1e-5
        assert np.allclose(_reference_test_trsv(m, rp, ci, v, b, upper, True), x, rtol=2e-4, atol=1e-5)
        xe = oracle.trsv(m, drp, dci, dv, b, upper=upper, unit=False)
        assert np.array_equal(xe, g[f"x_explicit_{name}_{key}"])
        assert np.allclose(_reference_test_trsv(m, drp, dci, dv, b, upper, False), xe, rtol=2e-4, atol=1e-5)
        xs = oracle.trsv(m, drp, dci, dv, b, upper=upper, unit=False, alpha_b=1.2)
        assert np.array_equal(xs, g[f"x_explicit_scaled_{name}_{key}"])
    # the reference test itself: b = 0 gives x = 0 whatever x held (triangular_solve_test.cpp:70-88)
    assert not oracle.trsv(m, rp, ci, v, np.zeros(m, np.float32), unit=True, x0=np.ones(m, np.float32)).any()


def test_trsv_against_real_reference_when_present(oracle):
    """Bit-exact against the real spblas::triangular_solve for every type combination, both
    triangles, both diagonal modes, scaled(a) / scaled(b) — including the reference's use of
    the previous row's diagonal when a row stores none (random matrices have such rows)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(41)
    for (vt, it, ot) in [(np.float32, np.int32, np.int32), (np.float32, np.int32, np.int64),
                         (np.float64, np.int32, np.int32), (np.float64, np.int32, np.int64),
                         (np.float32, np.int64, np.int64)]:
        m = 211
        lens = rng.integers(0, 14, size=m)
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(ot)
        ci = rng.integers(0, m, size=int(rp[-1])).astype(it)
        v = (0.05 * rng.standard_normal(len(ci))).astype(vt)
        b = rng.standard_normal(m).astype(vt)
        for upper in (0, 1):
            for unit in (0, 1):
                for kw in ({}, {"alpha_b": 1.2}, {"alpha_a": 0.5}, {"alpha_a": -2.0, "alpha_b": 3.0}):
                    got = oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit, **kw)
                    want = oracle.trsv(m, rp, ci, v, b, upper=upper, unit=unit, impl="reference", **kw)
                    assert np.array_equal(got, want, equal_nan=True)


def test_trsv_levels(oracle):
    # 5-point Poisson on a g x g grid: the lower triangle has 2g - 1 levels (anti-diagonals)
    import torch
    from spblas_reference_b200 import generators as G
    g = 9
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cpu")
    lv = oracle.trsv_levels(shape[0], rp.numpy(), ci.numpy())
    assert lv.max() + 1 == 2 * g - 1
    assert np.array_equal(lv, (np.arange(g * g) // g) + (np.arange(g * g) % g))
    lu = oracle.trsv_levels(shape[0], rp.numpy(), ci.numpy(), upper=True)
    assert np.array_equal(lu, lv[::-1])
    # a diagonal matrix has one level; a bidiagonal one has m
    assert oracle.trsv_levels(4, [0, 1, 2, 3, 4], [0, 1, 2, 3]).max() == 0
    assert oracle.trsv_levels(4, [0, 1, 3, 5, 7], [0, 0, 1, 1, 2, 2, 3]).tolist() == [0, 1, 2, 3]


def test_fuzz_restatement_against_the_real_reference(oracle):
    """Property test (hypothesis): on arbitrary small structures — empty matrices, empty rows,
    unsorted and duplicate columns, one-row / one-column shapes — the C restatement and the
    real spblas::multiply agree bit for bit, for SpMV and SpMM, CSR and CSC, plain and scaled."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def case(draw):
        m = draw(st.integers(0, 12))
        n = draw(st.integers(1, 9))
        lens = [draw(st.integers(0, 7)) for _ in range(m)]
        seed = draw(st.integers(0, 2 ** 31 - 1))
        k = draw(st.integers(1, 5))
        vt = draw(st.sampled_from([np.float32, np.float64, np.int32]))
        return m, n, lens, seed, k, vt

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
    @given(case())
    def run(c):
        m, n, lens, seed, k, vt = c
        rng = np.random.default_rng(seed)
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(rp[-1])
        ci = rng.integers(0, n, size=nnz).astype(np.int32)         # unsorted, duplicates legal
        if vt == np.int32:
            v = rng.integers(-9, 9, size=nnz).astype(vt)
            x = rng.integers(-9, 9, size=n).astype(vt)
            B = rng.integers(-9, 9, size=(n, k)).astype(vt)
            xt = rng.integers(-9, 9, size=m).astype(vt)
            aa = 3
        else:
            v = rng.standard_normal(nnz).astype(vt)
            x = rng.standard_normal(n).astype(vt)
            B = rng.standard_normal((n, k)).astype(vt)
            xt = rng.standard_normal(m).astype(vt)
            aa = 0.75
        for kw in ({}, {"alpha_a": aa}):
            assert np.array_equal(oracle.spmv("csr", (m, n), rp, ci, v, x, **kw),
                                  oracle.spmv("csr", (m, n), rp, ci, v, x, impl="reference", **kw))
            assert np.array_equal(oracle.spmm("csr", (m, n), rp, ci, v, B, **kw),
                                  oracle.spmm("csr", (m, n), rp, ci, v, B, impl="reference", **kw))
        # the same arrays read column-major: the n x m transpose
        assert np.array_equal(oracle.spmv("csc", (n, m), rp, ci, v, xt),
                              oracle.spmv("csc", (n, m), rp, ci, v, xt, impl="reference"))

    run()


def test_fuzz_transpose_and_trsv_against_the_real_reference(oracle):
    """Property test (hypothesis): transpose (structure AND values) and triangular_solve (both
    triangles, both diagonal modes, rows with or without a stored diagonal, duplicates) agree
    bit for bit with the real reference on arbitrary small structures."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(1, 11), st.integers(1, 8), st.integers(0, 2 ** 31 - 1),
           st.sampled_from([np.float32, np.float64]), st.integers(0, 6))
    def run(m, n, seed, vt, maxlen):
        rng = np.random.default_rng(seed)
        lens = rng.integers(0, maxlen + 1, size=m)
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        nnz = int(rp[-1])
        ci = rng.integers(0, n, size=nnz).astype(np.int32)
        v = rng.standard_normal(nnz).astype(vt)
        for a, b in zip(oracle.transpose((m, n), rp, ci, v),
                        oracle.transpose((m, n), rp, ci, v, impl="reference")):
            assert a.dtype == b.dtype and np.array_equal(a, b)
        # a square matrix over the same rows for the solve
        cs = rng.integers(0, m, size=nnz).astype(np.int32)
        vs = (0.2 * rng.standard_normal(nnz)).astype(vt)
        b = rng.standard_normal(m).astype(vt)
        for upper in (0, 1):
            for unit in (0, 1):
                got = oracle.trsv(m, rp, cs, vs, b, upper=upper, unit=unit)
                want = oracle.trsv(m, rp, cs, vs, b, upper=upper, unit=unit, impl="reference")
                assert np.array_equal(got, want, equal_nan=True)

    run()
