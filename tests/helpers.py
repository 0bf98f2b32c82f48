"""Shared helpers for the GPU parity tests: move numpy fixtures to the device, call the
backend through its host API (which goes through the C ABI), and check results with the
north-star tolerance |dy_i| <= (len_i + 2) * eps * sum_j |alpha a_ij x_j|."""
import numpy as np
import torch

import spblas_reference_b200 as sb

EPS = {np.dtype(np.float32): 2.0 ** -23, np.dtype(np.float64): 2.0 ** -52}


def dev(a, device="cuda:0"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def csr_on_device(values, rowptr, colind, shape, device="cuda:0"):
    nnz = int(rowptr[-1] - rowptr[0]) if len(rowptr) else 0
    return sb.csr_view(dev(values, device), dev(rowptr, device), dev(colind, device), shape, nnz)


def csc_on_device(values, colptr, rowind, shape, device="cuda:0"):
    nnz = int(colptr[-1] - colptr[0]) if len(colptr) else 0
    return sb.csc_view(dev(values, device), dev(colptr, device), dev(rowind, device), shape, nnz)


def gpu_spmv(a_view, x, m, dtype, alpha_a=None, alpha_x=None, info=None, poison=True,
             execute=False):
    xd = dev(x)
    y = torch.full((m,), float("nan") if np.dtype(dtype).kind == "f" else 77,
                   dtype=xd.dtype, device=xd.device) if poison else torch.zeros(
                       m, dtype=xd.dtype, device=xd.device)
    a = sb.scaled(alpha_a, a_view) if alpha_a is not None else a_view
    xv = sb.scaled(alpha_x, xd) if alpha_x is not None else xd
    if info is None:
        sb.multiply(a, xv, y)
    elif execute:
        sb.multiply_execute(info, a, xv, y)
    else:
        sb.multiply(info, a, xv, y)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def gpu_spmm(a_view, B, m, alpha_a=None, alpha_b=None, info=None):
    Bd = dev(B)
    Cd = torch.full((m, B.shape[1]), float("nan") if B.dtype.kind == "f" else 77,
                    dtype=Bd.dtype, device=Bd.device)
    a = sb.scaled(alpha_a, a_view) if alpha_a is not None else a_view
    Bv = sb.scaled(alpha_b, Bd) if alpha_b is not None else Bd
    if info is None:
        sb.multiply(a, Bv, Cd)
    else:
        sb.multiply(info, a, Bv, Cd)
    torch.cuda.synchronize()
    return Cd.cpu().numpy()


def assert_rows_within_bound(y_gpu, y_ref, rowptr, bound_sum, what=""):
    """|y_gpu - y_ref| <= (len_i + 2) * eps * s_i with s_i = sum |alpha a x| (float64).
    y may be (m,) or (m, k) with bound_sum of the same shape."""
    y_gpu, y_ref = np.asarray(y_gpu), np.asarray(y_ref)
    if y_ref.dtype.kind != "f":
        assert np.array_equal(y_gpu, y_ref), f"{what}: integer results must be exact"
        return 0.0
    eps = EPS[y_ref.dtype]
    lens = np.diff(np.asarray(rowptr).astype(np.int64)).astype(np.float64)
    if y_ref.ndim == 2:
        lens = lens[:, None]
    tol = (lens + 2.0) * eps * bound_sum
    err = np.abs(y_gpu.astype(np.float64) - y_ref.astype(np.float64))
    assert np.isfinite(y_gpu).all(), f"{what}: non-finite output"
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} entries exceed the per-row bound; worst "
                           f"err/tol = {np.max(err[bad] / np.maximum(tol[bad], 1e-300)):.3g}")
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(tol > 0, err / tol, 0.0)
    return float(ratio.max()) if ratio.size else 0.0


def spmm_bound(rowptr, colind, values, B, alpha=1.0):
    """s_ij = sum_k |alpha a_ik B_kj| in float64."""
    m = len(rowptr) - 1
    rows = np.repeat(np.arange(m), np.diff(np.asarray(rowptr).astype(np.int64)))
    out = np.zeros((m, B.shape[1]), dtype=np.float64)
    contrib = np.abs(alpha * values.astype(np.float64))[:, None] * np.abs(
        B.astype(np.float64))[np.asarray(colind).astype(np.int64)]
    np.add.at(out, rows, contrib)
    return out
