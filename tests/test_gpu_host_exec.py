"""GPU parity tests for the host-buffer execute (spblas_b200_spmv_host, csrc/host_exec.cu):
upload of x, the SpMV kernels and the download of y pipelined chunk by chunk.  The result
must be BIT-IDENTICAL to the device-vector execute (same kernels, same tiles) and within
the north-star bound of the oracle, whatever the chunk count, kernel variant, row-length
mix (rows spanning several tiles and chunk boundaries), storage format or column spread."""
import os
import zlib

import numpy as np
import pytest
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from helpers import assert_rows_within_bound, csc_on_device, csr_on_device, dev
from test_gpu_spmv import _random_csr

pytestmark = pytest.mark.gpu


def _host_vs_device(a, x, m, vt, alpha=None):
    xd = dev(x)
    av = sb.scaled(alpha, a) if alpha is not None else a
    y_dev = torch.full((m,), 7, dtype=xd.dtype, device="cuda")
    info = sb.multiply_inspect(a, xd, y_dev)
    sb.multiply_execute(info, av, xd, y_dev)
    xh = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
    yh = torch.full((m,), 9, dtype=xd.dtype).pin_memory()
    for _ in range(2):                       # twice: staging buffers and events are reused
        yh.fill_(9)
        sb.multiply_execute_host(info, av, xh, yh)
        torch.cuda.synchronize()
        assert np.array_equal(yh.numpy(), y_dev.cpu().numpy(), equal_nan=True)
    info.close()
    return yh.numpy().copy()


@pytest.mark.parametrize("chunks", ["1", "3", "16", "64"])
@pytest.mark.parametrize("variant", ["0", "1", "2"])
@pytest.mark.parametrize("kind", ["short", "mixed", "hub", "empty"])
def test_host_execute_bit_identical(cuda, oracle, monkeypatch, kind, variant, chunks):
    monkeypatch.setenv("SPBLAS_B200_HOST_CHUNKS", chunks)
    monkeypatch.setenv("SPBLAS_B200_SPMV_VARIANT", variant)
    monkeypatch.setenv("SPBLAS_B200_TILE_ITEMS", "512")      # many tiles -> many chunk cuts
    vt = np.float64 if kind != "mixed" else np.float32
    rng = np.random.default_rng(zlib.crc32(f"host{kind}".encode()))
    m, n = 6001, 3777
    v, rp, ci, x = _random_csr(rng, m, n, kind, vt, np.int32, np.int32)
    a = csr_on_device(v, rp, ci, (m, n))
    y = _host_vs_device(a, x, m, vt, alpha=0.75)
    y_ref = oracle.spmv("csr", (m, n), rp, ci, v, x, alpha_a=0.75)
    assert_rows_within_bound(y, y_ref, rp, oracle.abs_rowsum(rp, ci, v, x, 0.75), f"host {kind}")


def test_host_execute_banded_progressive_upload(cuda, oracle, monkeypatch):
    """Poisson stencil: chunk c only needs x up to its last row + one grid line, so the
    upload is progressive; int32 scalars make the comparison with the oracle exact."""
    monkeypatch.setenv("SPBLAS_B200_HOST_CHUNKS", "8")
    g = 96
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cuda:0")
    vh, rph, cih = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy()
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x = np.random.default_rng(5).standard_normal(shape[1])
    y = _host_vs_device(a, x, shape[0], np.float64, alpha=0.125)
    y_ref = oracle.spmv("csr", shape, rph, cih, vh, x, alpha_a=0.125)
    assert_rows_within_bound(y, y_ref, rph, oracle.abs_rowsum(rph, cih, vh, x, 0.125), "host banded")
    ai = sb.csr_view(v.to(torch.int32), rp, ci, shape, int(ci.numel()))
    xi = np.random.default_rng(6).integers(-9, 10, size=shape[1]).astype(np.int32)
    yi = _host_vs_device(ai, xi, shape[0], np.int32)
    assert np.array_equal(yi, oracle.spmv("csr", shape, rph, cih, vh.astype(np.int32), xi))


def test_host_execute_csc_and_int64_offsets(cuda, oracle, monkeypatch):
    monkeypatch.setenv("SPBLAS_B200_HOST_CHUNKS", "5")
    rng = np.random.default_rng(8)
    m, n = 2500, 3100
    # a CSC matrix: "rows" of the generator are the columns of A
    v, cp, ri, _ = _random_csr(rng, n, m, "mixed", np.float64, np.int32, np.int64)
    a = csc_on_device(v, cp, ri, (m, n))
    x = rng.standard_normal(n)
    y = _host_vs_device(a, x, m, np.float64)
    y_ref = oracle.spmv("csc", (m, n), cp, ri, v, x)
    t_rp, t_ci, perm = oracle.csc_row_major_image((m, n), cp, ri)
    assert_rows_within_bound(y, y_ref, t_rp, oracle.abs_rowsum(t_rp, t_ci, v[perm], x), "host csc")


def test_host_execute_errors(cuda):
    v, rp, ci, shape = G.poisson2d_csr(8, torch.float32, "cuda:0")
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    info = sb.operation_info_t()
    xh, yh = torch.ones(shape[1]), torch.empty(shape[0])
    with pytest.raises(ValueError):                      # std::invalid_argument
        sb.multiply_execute_host(info, a, torch.ones(shape[1] + 1), yh)
    with pytest.raises(RuntimeError):
        sb.multiply_execute_host(info, a, xh.cuda(), yh)
    sb.multiply_execute_host(info, a, xh, yh)            # pageable memory works too (synchronous copies)
    torch.cuda.synchronize()
    y = torch.empty(shape[0], device="cuda")
    sb.multiply(a, xh.cuda(), y)
    assert torch.equal(y.cpu(), yh)
    info.close()
