"""GPU parity tests for the hub-column SpMV variant (csrc/hub.cu + spmv_hub_stream_kernel,
variant 3): the inspect-phase structures bit-exactly against the oracle's definition
(oracle.hub_columns), the product bit-identically against the warp-stream kernel (same
arithmetic in the same order) and within the north-star bound against the reference's
multiply, on skewed / uniform / hub-row matrices, every scalar type, int64 offsets, row-block
shards with unaligned bases, CSC operands, and the automatic choice.

(The file name sorts last on purpose: a new kernel is tested after everything that was
already green.)"""
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from helpers import assert_rows_within_bound, csc_on_device, csr_on_device, dev

pytestmark = pytest.mark.gpu


def _skewed_csr(rng, m, n, lens, vt, ot=np.int32, power=4.0):
    """columns drawn with density ~ c^(1/power - 1): a few columns take most references"""
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(ot)
    nnz = int(rp[-1])
    ci = np.minimum((n * rng.random(nnz) ** power).astype(np.int64), n - 1).astype(np.int32)
    if vt == np.int32:
        v = rng.integers(-9, 10, size=nnz).astype(vt)
        x = rng.integers(-9, 10, size=n).astype(vt)
    else:
        v = rng.standard_normal(nnz).astype(vt)
        x = rng.standard_normal(n).astype(vt)
    return v, rp, ci, x


def _lens(rng, m, kind):
    if kind == "short":
        return rng.integers(0, 12, size=m)
    if kind == "hubrow":
        lens = rng.integers(0, 5, size=m)
        lens[m // 3], lens[m - 1], lens[0] = 9000, 2500, 2049
        return lens
    if kind == "long":
        return rng.integers(100, 300, size=m)
    raise ValueError(kind)


def _run(a, xd, m, variant, hub=None, alpha=None):
    y = torch.full((m,), float("nan") if xd.dtype.is_floating_point else 77, dtype=xd.dtype,
                   device=xd.device)
    info = sb.multiply_inspect(a, xd, y)
    if hub is not None:
        info.set_hub(True, *hub)
    if variant is not None:
        info.force_spmv_variant(variant)
    sb.multiply_execute(info, sb.scaled(alpha, a) if alpha is not None else a, xd, y)
    torch.cuda.synchronize()
    return y, info


@pytest.mark.parametrize("kind", ["short", "hubrow", "long"])
@pytest.mark.parametrize("types", [(np.float32, np.int32), (np.float64, np.int32),
                                   (np.int32, np.int32), (np.float64, np.int64)])
def test_hub_structures_and_product(cuda, oracle, kind, types):
    vt, ot = types
    rng = np.random.default_rng(zlib.crc32(f"hub{kind}{vt.__name__}{ot.__name__}".encode()))
    m, n = 5003, 2777
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, kind), vt, ot)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    alpha = 3 if vt == np.int32 else 0.75
    y_ws, i_ws = _run(a, xd, m, 2, alpha=alpha)
    y_hub, i_hub = _run(a, xd, m, 3, hub=(64, 3), alpha=alpha)
    assert i_ws.spmv_variant == 2 and i_hub.spmv_variant == 3
    # the inspect-phase structures, bit for bit
    hubs, refs, enc = oracle.hub_columns(ci, n, 64, 3)
    assert i_hub.hub_count == len(hubs) and len(hubs) == 64
    assert np.array_equal(i_hub.hub_cols, hubs)
    assert i_hub.hub_refs == refs and refs > len(ci) // 4
    assert np.array_equal(i_hub.hub_colind, enc)
    # the same arithmetic in the same order as the warp-stream kernel
    assert torch.equal(y_ws, y_hub)
    y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=alpha)
    bound = None if vt == np.int32 else oracle.abs_rowsum(rp, ci, v, x, alpha)
    assert_rows_within_bound(y_hub.cpu().numpy(), y_ref, rp, bound, f"hub {kind} {vt.__name__}")
    # values may change between executes; the structure copy does not depend on them
    a.values.mul_(2)
    y2 = torch.empty_like(y_hub)
    sb.multiply_execute(i_hub, sb.scaled(alpha, a), xd, y2)
    y2_ref = oracle.spmv("csr", (m, n), rp, ci, (v * 2).astype(vt), x, alpha_a=alpha)
    assert_rows_within_bound(y2.cpu().numpy(), y2_ref, rp,
                             None if bound is None else 2 * bound, "hub after a change of values")
    i_ws.close()
    i_hub.close()


@pytest.mark.parametrize("kind", ["short", "hubrow", "long"])
@pytest.mark.parametrize("types", [(np.float32, np.int32), (np.float64, np.int64), (np.int32, np.int32)])
def test_global_hub_table_structures_and_product(cuda, oracle, kind, types):
    """Variant 4: the hub table in global memory, hottest column first (for x larger than L2).
    Table, encoded colind and reference count against the oracle's definition; y bit-identical
    to the warp-stream kernel's."""
    vt, ot = types
    rng = np.random.default_rng(zlib.crc32(f"hubg{kind}{vt.__name__}{ot.__name__}".encode()))
    m, n = 6007, 4099
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, kind), vt, ot)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    alpha = 3 if vt == np.int32 else 0.75
    y_ws, i_ws = _run(a, xd, m, 2, alpha=alpha)
    y_hub, i_hub = _run(a, xd, m, 4, hub=(300, 3), alpha=alpha)
    hubs, refs, enc = oracle.hub_columns(ci, n, 300, 3, by_popularity=True)
    assert i_hub.spmv_variant == 4
    assert i_hub.hub_count == len(hubs) and i_hub.hub_refs == refs
    assert np.array_equal(i_hub.hub_cols, hubs)
    assert np.array_equal(i_hub.hub_colind, enc)
    assert torch.equal(y_ws, y_hub)
    y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=alpha)
    bound = None if vt == np.int32 else oracle.abs_rowsum(rp, ci, v, x, alpha)
    assert_rows_within_bound(y_hub.cpu().numpy(), y_ref, rp, bound, f"global hub {kind}")
    # a second product with another x: the table is refilled by every product
    x2 = dev((x * 2).astype(vt))
    y2 = torch.empty_like(y_hub)
    sb.multiply_execute(i_hub, sb.scaled(alpha, a), x2, y2)
    y2_ws = torch.empty_like(y_hub)
    sb.multiply_execute(i_ws, sb.scaled(alpha, a), x2, y2_ws)
    assert torch.equal(y2, y2_ws)
    i_ws.close()
    i_hub.close()


def test_global_hub_table_is_not_chosen_for_a_small_x(cuda, oracle):
    """matrix_opt picks the global table only when x exceeds 3/4 of L2; a small x keeps the
    shared-memory table or the plain walk."""
    rng = np.random.default_rng(5)
    m, n = 4001, 3001
    v, rp, ci, x = _skewed_csr(rng, m, n, _lens(rng, m, "hubrow"), np.float64)
    a = sb.matrix_opt(csr_on_device(v, rp, ci, (m, n)))
    xd = dev(x)
    y = torch.empty(m, dtype=torch.float64, device="cuda")
    info = sb.multiply_inspect(a, xd, y)
    sb.multiply_execute(info, a, xd, y)
    assert info.spmv_variant in (2, 3)
    info.close()


@pytest.mark.parametrize("r0", [777, 778, 779, 780])
def test_hub_on_a_row_block_with_an_unaligned_base(cuda, oracle, r0):
    """A shard keeps the global rowptr base; rp[r0] % 4 takes every value over the four
    starts, so the plan's copy of colind starts 0..3 entries before the shard's first."""
    rng = np.random.default_rng(5)
    m, n, r1 = 5003, 2777, 4100
    lens = rng.integers(0, 12, size=m)
    lens[:r0] = 3                                     # rp[r0] = 3 r0: all residues mod 4
    v, rp, ci, x = _skewed_csr(rng, m, n, lens, np.float64)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    blk = sb.csr_view(a.values, a.rowptr[r0:r1 + 1], a.colind, (r1 - r0, n),
                      int(rp[r1] - rp[r0]))
    y_ws, i_ws = _run(blk, xd, r1 - r0, 2)
    y_hub, i_hub = _run(blk, xd, r1 - r0, 3, hub=(100, 2))
    assert i_hub.spmv_variant == 3
    hubs, refs, enc = oracle.hub_columns(ci[rp[r0]:rp[r1]], n, 100, 2)
    assert np.array_equal(i_hub.hub_cols, hubs) and i_hub.hub_refs == refs
    assert np.array_equal(i_hub.hub_colind, enc)
    assert torch.equal(y_ws, y_hub)
    y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x)[r0:r1]
    assert_rows_within_bound(y_hub.cpu().numpy(), y_ref, rp[r0:r1 + 1],
                             oracle.abs_rowsum(rp, ci, v, x)[r0:r1], f"hub shard base {rp[r0] % 4}")
    i_ws.close()
    i_hub.close()


def test_hub_on_a_csc_operand_and_cached_values(cuda, oracle):
    """CSC: the hub analysis runs on the plan's row-major image; values through the
    permutation, or cached in image order under matrix_opt."""
    rng = np.random.default_rng(17)
    m, n = 2203, 3001                                           # A is m x n, stored by columns
    lens = rng.integers(0, 9, size=n)
    lens[rng.choice(n, size=60, replace=False)] = rng.integers(300, 900, size=60)  # popular columns
    cp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ri = rng.integers(0, m, size=int(cp[-1])).astype(np.int32)
    v = rng.standard_normal(len(ri))
    x = rng.standard_normal(n)
    ac = csc_on_device(v, cp, ri, (m, n))
    xd = dev(x)
    y_ws, i_ws = _run(ac, xd, m, 2)
    y_hub, i_hub = _run(ac, xd, m, 3, hub=(128, 1))
    assert i_hub.spmv_variant == 3
    t_rp, t_ci, perm = oracle.csc_row_major_image((m, n), cp, ri)
    hubs, refs, enc = oracle.hub_columns(t_ci, n, 128, 1)
    assert np.array_equal(i_hub.hub_cols, hubs) and np.array_equal(i_hub.hub_colind, enc)
    assert torch.equal(y_ws, y_hub)
    assert_rows_within_bound(y_hub.cpu().numpy(), oracle.spmv("csc", (m, n), cp, ri, v, x), t_rp,
                             oracle.abs_rowsum(t_rp, t_ci, v[perm], x), "hub csc")
    aopt = sb.matrix_opt(ac)
    y_opt, i_opt = _run(aopt, xd, m, 3, hub=(128, 1))
    assert i_opt.spmv_variant == 3 and torch.equal(y_opt, y_hub)
    for i in (i_ws, i_hub, i_opt):
        i.close()


def test_hub_automatic_choice(cuda, oracle):
    """Enabled, the hub variant replaces the warp-stream kernel only where it pays: skewed
    columns yes; uniform columns no (and the copy of colind is given back); a stencil keeps
    the pipelined kernel; the no-info overload never analyses; int64 indices fall back."""
    rng = np.random.default_rng(23)
    m, n = 20011, 9973
    lens = rng.integers(0, 24, size=m)
    v, rp, ci, x = _skewed_csr(rng, m, n, lens, np.float32)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    y_ws, i_ws = _run(a, xd, m, None)
    assert i_ws.spmv_variant == 2 and i_ws.hub_count == 0
    y_hub, i_hub = _run(a, xd, m, None, hub=(0, 4))
    assert i_hub.spmv_variant == 3 and 0 < i_hub.hub_count <= 32768
    hubs, refs, _ = oracle.hub_columns(ci, n, 32768, 4)
    assert i_hub.hub_count == len(hubs) and i_hub.hub_refs == refs
    assert torch.equal(y_ws, y_hub)
    # the no-info overload (a light inspect per call, never analysed): the plain walk's result
    y1 = torch.empty_like(y_ws)
    sb.multiply(a, xd, y1)
    assert torch.equal(y1, y_ws)
    # uniform columns at the default threshold (2 x SM count references): no hubs
    ci_u = rng.integers(0, n, size=len(ci)).astype(np.int32)
    au = csr_on_device(v, rp, ci_u, (m, n))
    y_u, i_u = _run(au, xd, m, None, hub=(0, 0))
    assert i_u.spmv_variant == 2 and i_u.hub_count == 0
    assert_rows_within_bound(y_u.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci_u, v, x), rp,
                             oracle.abs_rowsum(rp, ci_u, v, x), "uniform columns, hub offered")
    # int64 column indices: the walk stays on the plain kernel even when forced
    a64 = sb.csr_view(a.values, a.rowptr.to(torch.int64), a.colind.to(torch.int64), (m, n), len(ci))
    y64, i64 = _run(a64, xd, m, 3, hub=(0, 4))
    assert i64.spmv_variant == 2 and torch.equal(y64, y_ws)
    # a stencil keeps the TMA pipeline
    from spblas_reference_b200 import generators as G
    vs, rps, cis, shape = G.poisson2d_csr(1024, torch.float64, cuda)   # (a small grid's tiles mix 4- and 5-entry rows)
    st = sb.csr_view(vs, rps, cis, shape, int(cis.numel()))
    xs = torch.ones(shape[1], dtype=torch.float64, device=cuda)
    _, i_st = _run(st, xs, shape[0], None, hub=(0, 1))
    assert i_st.spmv_variant == 1
    for i in (i_ws, i_hub, i_u, i64, i_st):
        i.close()


def test_hub_default_capacity_and_value_width_switch(cuda, oracle, monkeypatch):
    """Default limits on a matrix large enough to fill the table: 32768 columns for 4-byte
    values (a 164 KB carve-out), 8192 for 8-byte ones (131 KB); L1 keeps the rest; the table is
    rebuilt when the value width changes; an explicit max_cols may go up to what the SM's
    shared memory holds (49152 / 20480)."""
    # (the stream length follows the number of resident warps, which differs between the two
    # kernels; pinned so that both cut the rows at the same places and agree bit for bit)
    monkeypatch.setenv("SPBLAS_B200_WS_ITEMS", "1024")
    rng = np.random.default_rng(29)
    m, n = 120_000, 400_000
    lens = np.full(m, 40)
    v, rp, ci, x = _skewed_csr(rng, m, n, lens, np.float32, power=3.0)
    a32 = csr_on_device(v, rp, ci, (m, n))
    x32 = dev(x)
    y_ws, i_ws = _run(a32, x32, m, 2)
    y_hub, info = _run(a32, x32, m, 3, hub=(0, 8))
    hubs, refs, enc = oracle.hub_columns(ci, n, 32768, 8)
    assert info.hub_count == len(hubs) == 32768 and info.hub_refs == refs
    assert np.array_equal(info.hub_cols, hubs) and np.array_equal(info.hub_colind, enc)
    assert torch.equal(y_ws, y_hub)
    # the same plan, fp64 values: smaller table
    a64 = sb.csr_view(a32.values.double(), a32.rowptr, a32.colind, (m, n), len(ci))
    x64 = x32.double()
    y64 = torch.empty(m, dtype=torch.float64, device=cuda)
    sb.multiply_execute(info, a64, x64, y64)
    torch.cuda.synchronize()
    hubs64, refs64, enc64 = oracle.hub_columns(ci, n, 8192, 8)
    assert info.spmv_variant == 3 and info.hub_count == 8192 and info.hub_refs == refs64
    assert np.array_equal(info.hub_cols, hubs64) and np.array_equal(info.hub_colind, enc64)
    y_ref = oracle.spmv("csr", (m, n), rp, ci, v.astype(np.float64), x.astype(np.float64))
    assert_rows_within_bound(y64.cpu().numpy(), y_ref, rp,
                             oracle.abs_rowsum(rp, ci, v.astype(np.float64), x.astype(np.float64)),
                             "hub fp64 after fp32")
    # an explicit request is clamped to the hardware limit, not to the default
    y_max, i_max = _run(a32, x32, m, 3, hub=(1 << 20, 8))
    assert i_max.hub_count == 49152 and torch.equal(y_max, y_ws)
    i_max.close()
    i_ws.close()
    info.close()


def test_hub_plan_serves_the_host_buffer_execute(cuda, oracle):
    """multiply_execute_host on a hub-enabled plan: the chunked launches take the plain walk
    (bit-identical), and the device-vector execute afterwards is the hub kernel again."""
    rng = np.random.default_rng(31)
    m, n = 30011, 9973
    v, rp, ci, x = _skewed_csr(rng, m, n, rng.integers(0, 24, size=m), np.float64)
    a = csr_on_device(v, rp, ci, (m, n))
    xd = dev(x)
    y_hub, info = _run(a, xd, m, None, hub=(0, 4))
    assert info.spmv_variant == 3
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.empty(m, dtype=torch.float64).pin_memory()
    sb.multiply_execute_host(info, a, xh, yh)
    torch.cuda.synchronize()
    assert info.spmv_variant == 2
    assert torch.equal(yh, y_hub.cpu())
    y3 = torch.empty_like(y_hub)
    sb.multiply_execute(info, a, xd, y3)
    torch.cuda.synchronize()
    assert info.spmv_variant == 3 and torch.equal(y3, y_hub)
    info.close()


def test_hub_rmat_reduced(cuda, oracle, monkeypatch):
    """C4's shape at scale 18: R-MAT columns are what the variant is for."""
    monkeypatch.setenv("SPBLAS_B200_WS_ITEMS", "1024")   # same stream cuts for both kernels
    from spblas_reference_b200 import generators as G
    v, rp, ci, shape = G.rmat_csr(18, 16, seed=24, dtype=torch.float32, device=cuda)
    m, n = shape
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = G.dense_uniform((n,), 5, torch.float32, cuda)
    y_ws, i_ws = _run(a, x, m, 2)
    y_hub, i_hub = _run(a, x, m, None, hub=(0, 0))
    assert i_hub.spmv_variant == 3
    assert 3 * i_hub.hub_refs > int(ci.numel())          # > 1/3 of the gathers leave the L2 port
    assert torch.equal(y_ws, y_hub)
    vh, rph, cih, xh = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy(), x.cpu().numpy()
    y_ref = oracle.spmv("csr", shape, rph, cih, vh, xh)
    assert_rows_within_bound(y_hub.cpu().numpy(), y_ref, rph, y_ref.astype(np.float64), "hub R-MAT 18")
    i_ws.close()
    i_hub.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
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
        y, info = _run(a, xd, m, variant, hub=(16, 1) if variant >= 3 else None, alpha=2)
        assert np.array_equal(y.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=2))
        info.close()
        if m >= 4:                                   # rows [r0, r1) with the global base
            r0, r1 = m // 4, m - m // 4
            blk = sb.csr_view(a.values, a.rowptr[r0:r1 + 1], a.colind, (r1 - r0, n),
                              int(rp[r1] - rp[r0]))
            yb, ib = _run(blk, xd, r1 - r0, variant, hub=(16, 1) if variant >= 3 else None)
            assert np.array_equal(yb.cpu().numpy(), oracle.spmv("csr", (m, n), rp, ci, v, x)[r0:r1])
            ib.close()

    run()
