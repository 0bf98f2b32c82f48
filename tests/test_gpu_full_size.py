"""Parity at BASELINE.json's full sizes, against the REAL reference's CPU multiply (oracle/_ref
when it travelled with the snapshot, else the oracle port) under the north-star bound — the
same check bench.py attaches to every timed config (bench_configs.check_spmv / check_spmm),
here as tests: C2 (Poisson 4096^2, fp64, scaled(1/8, a)), C4 (R-MAT scale 24, fp32, through
matrix_opt AND plain: bit-identical), C1 through both overloads, C3 k = 32 on the full 2M x 2M
matrix (reference on a 200k-row block).  C5's 2^31 entries (29 GB) are checked by the bench
line itself (`configs.c5.parity`, including rows whose offsets exceed 2^31 - 1)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_c2_full_size(cuda, oracle):
    from bench_configs import check_spmv
    v, rp, ci, shape = G.poisson2d_csr(4096, torch.float64, DEV)
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = G.dense_uniform((shape[1],), 1, torch.float64, DEV)
    y = torch.full((shape[0],), float("nan"), dtype=torch.float64, device=DEV)
    info = sb.multiply_inspect(a, x, y)
    sb.multiply_execute(info, sb.scaled(0.125, a), x, y)
    assert info.spmv_variant == 1
    _, parity = check_spmv(rp, ci, v, x, 0.125, y, (0, shape[0]), shape[1], "C2")
    assert parity["pass"] and parity["rows_checked"] == 4096 * 4096, parity
    info.close()


def test_c4_full_size_plain_and_matrix_opt(cuda, oracle):
    from bench_configs import check_spmv
    v, rp, ci, shape = G.rmat_csr(24, 16, seed=24, dtype=torch.float32, device=DEV)
    m, n = shape
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = G.dense_uniform((n,), 5, torch.float32, DEV)
    y = torch.full((m,), float("nan"), device=DEV)
    info = sb.multiply_inspect(a, x, y)
    sb.multiply_execute(info, a, x, y)
    _, parity = check_spmv(rp, ci, v, x, None, y, (0, m), n, "C4")
    assert parity["pass"] and parity["rows_checked"] == m, parity
    y2 = torch.full((m,), float("nan"), device=DEV)
    a_opt = sb.matrix_opt(a)
    info2 = sb.multiply_inspect(a_opt, x, y2)
    sb.multiply_execute(info2, a_opt, x, y2)
    assert info.spmv_variant == 2 and info2.spmv_variant == 3
    assert torch.equal(y, y2)
    info.close()
    info2.close()


def test_c1_full_size_both_overloads(cuda, oracle):
    from bench_configs import check_spmv
    m = n = 1_000_000
    v, rp, ci, shape = G.uniform_random_csr(m, n, 10, seed=0, dtype=torch.float32, device=DEV)
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = G.dense_uniform((n,), 100, torch.float32, DEV)
    y = torch.full((m,), float("nan"), device=DEV)
    info = sb.multiply_inspect(a, x, y)
    sb.multiply_execute(info, sb.scaled(1.2, a), x, y)
    _, parity = check_spmv(rp, ci, v, x, 1.2, y, (0, m), n, "C1")
    assert parity["pass"], parity
    for _ in range(3):                      # no info: first call inspects, the others reuse
        y2 = torch.full((m,), float("nan"), device=DEV)
        sb.multiply(sb.scaled(1.2, a), x, y2)
        assert torch.equal(y, y2)
    info.close()


@pytest.mark.parametrize("k", [32, 128])
def test_c3_full_size(cuda, oracle, k):
    from bench_configs import check_spmm
    m = n = 2_000_000
    v, rp, ci, shape = G.uniform_random_csr(m, n, 16, seed=3, dtype=torch.float32, device=DEV)
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    B = G.dense_uniform((n, k), 4, torch.float32, DEV)
    C = torch.full((m, k), float("nan"), device=DEV)
    sb.multiply(a, B, C)
    assert bool(torch.isfinite(C).all())
    rows = (900_000, 1_100_000) if k == 32 else (1_950_000, 2_000_000)   # a middle block / the last rows
    _, parity = check_spmm(rp, ci, v, B, None, C[rows[0]:rows[1]], rows, n, f"C3 k={k}", compact=True)
    assert parity["pass"], parity
