"""Synthetic matrices of the shapes BASELINE.json names, generated with
counter-based hashing so that any device (CPU for the oracle, CUDA for the backend) and
any row block of a sharded run produces identical entries.  Data generation is harness
plumbing (torch ops), not part of the measured path.

  C1  uniform_random_csr(m, n, 10, fp32)         examples/simple_spmv-style random CSR
  C2  poisson2d_csr(4096, fp64)                  5-point stencil, rows in ascending column order
  C3  uniform_random_csr(2M, 2M, 16, fp32)       SpMM operand
  C4  rmat_csr(24, 16, fp32)                     power-law rows (hubs), duplicates kept
  C5  rmat_csr(27, 16, fp64, int64 offsets)      generated per row block
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

_MASK63 = (1 << 63) - 1


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right of int64 (torch's >> is arithmetic)."""
    return (x >> s) & ((1 << (64 - s)) - 1)


def _c64(v: int) -> int:
    """Python int -> the int64 with the same low 64 bits."""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's-complement wraparound)."""
    x = x + _c64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _c64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _c64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def hash_uniform(seed: int, idx: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """U[0,1) from (seed, idx): top 53 (fp64) or 24 (fp32) bits of splitmix64."""
    h = splitmix64(idx + _c64(seed * 0x632BE59BD9B4E019))
    if dtype == torch.float32:
        return (_lsr(h, 40).to(torch.float32)) * (1.0 / (1 << 24))
    return (_lsr(h, 11).to(torch.float64)) * (1.0 / (1 << 53))


def dense_uniform(shape, seed: int, dtype, device) -> torch.Tensor:
    n = 1
    for s in shape:
        n *= int(s)
    idx = torch.arange(n, dtype=torch.int64, device=device)
    if dtype == torch.int32:
        return (_lsr(splitmix64(idx + _c64(seed * 0x632BE59BD9B4E019)), 60) - 8).to(torch.int32).reshape(shape)
    return hash_uniform(seed, idx, dtype).reshape(shape)


def dense_uniform_rows(rows: int, cols: int, seed: int, dtype, device,
                       chunk_elems: int = 1 << 26) -> torch.Tensor:
    """dense_uniform((rows, cols), ...) filled in row chunks: the same values, with
    temporaries of chunk_elems elements instead of the whole operand's size (a replicated
    SpMM operand of tens of GB must not triple its footprint while it is generated)."""
    out = torch.empty((rows, cols), dtype=dtype, device=device)
    step = max(1, chunk_elems // max(cols, 1))
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        idx = torch.arange(r0 * cols, r1 * cols, dtype=torch.int64, device=device)
        if dtype == torch.int32:
            blk = (_lsr(splitmix64(idx + _c64(seed * 0x632BE59BD9B4E019)), 60) - 8).to(torch.int32)
        else:
            blk = hash_uniform(seed, idx, dtype)
        out[r0:r1] = blk.reshape(r1 - r0, cols)
        del idx, blk
    return out


# -------------------------------------------------------------------------------------
def poisson2d_csr(g: int, dtype=torch.float64, device="cpu", row_begin: int = 0,
                  row_end: Optional[int] = None, off_dtype=torch.int32,
                  gi: Optional[int] = None):
    """5-point Poisson stencil on a gi x g grid (gi = g by default; weak-scaled multi-GPU
    runs stack one g x g grid per GPU: gi = world * g), rows [row_begin, row_end) of the
    n = gi*g matrix: (-1, -1, 4, -1, -1) in ascending column order, boundary rows
    truncated.  Returns (values, rowptr, colind, (rows, n)); rowptr is rebased to 0."""
    gi = g if gi is None else gi
    n = gi * g
    row_end = n if row_end is None else row_end
    r = torch.arange(row_begin, row_end, dtype=torch.int64, device=device)
    i, j = r // g, r % g
    cols = torch.stack([r - g, r - 1, r, r + 1, r + g], dim=1)
    valid = torch.stack([i > 0, j > 0, torch.ones_like(i, dtype=torch.bool), j < g - 1,
                         i < gi - 1], dim=1)
    vals = torch.tensor([-1, -1, 4, -1, -1], dtype=dtype, device=device).expand(len(r), 5)
    counts = valid.sum(dim=1)
    rowptr = torch.zeros(len(r) + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    colind = cols[valid].to(torch.int32)
    values = vals[valid].contiguous()
    return values, rowptr.to(off_dtype), colind, (int(row_end - row_begin), n)


def uniform_random_csr(m: int, n: int, per_row: int, seed: int, dtype=torch.float32,
                       device="cpu", row_begin: int = 0, row_end: Optional[int] = None,
                       off_dtype=torch.int32):
    """Exactly `per_row` entries per row, columns = hash(seed, row, slot) mod n
    (unsorted; the rare duplicate is kept — duplicates accumulate),
    values U[0,1)."""
    row_end = m if row_end is None else row_end
    rows = row_end - row_begin
    e = torch.arange(row_begin * per_row, row_end * per_row, dtype=torch.int64, device=device)
    h = splitmix64(e + _c64(seed * 0x632BE59BD9B4E019))
    colind = ((h & _MASK63) % n).to(torch.int32)
    if dtype == torch.int32:
        values = (_lsr(splitmix64(h), 60) - 8).to(torch.int32)
    else:
        values = hash_uniform(seed + 1, e, dtype)
    rowptr = (torch.arange(rows + 1, dtype=torch.int64, device=device) * per_row).to(off_dtype)
    return values, rowptr, colind, (rows, n)


def rmat_edges(scale: int, edge_begin: int, edge_end: int, seed: int, device,
               abcd=(0.57, 0.19, 0.19, 0.05)) -> Tuple[torch.Tensor, torch.Tensor]:
    """R-MAT edges [edge_begin, edge_end): one hash per (edge, level) picks the quadrant."""
    a, b, c, _ = abcd
    e = torch.arange(edge_begin, edge_end, dtype=torch.int64, device=device)
    row = torch.zeros_like(e)
    col = torch.zeros_like(e)
    for level in range(scale):
        u = hash_uniform(seed * 64 + level + 1, e, torch.float64)
        right = ((u >= a) & (u < a + b)) | (u >= a + b + c)   # quadrants b, d -> column bit
        down = u >= a + b                                      # quadrants c, d -> row bit
        row = row * 2 + down.to(torch.int64)
        col = col * 2 + right.to(torch.int64)
    return row, col


def rmat_degrees(scale: int, edge_factor: int, seed: int, device, chunk_edges: int = 1 << 26):
    """Row degrees of the R-MAT graph (one pass over the edges, no storage): what a rank
    needs to cut nnz-balanced row blocks before it generates its own block."""
    n = 1 << scale
    total = edge_factor * n
    deg = torch.zeros(n, dtype=torch.int64, device=device)
    for b in range(0, total, chunk_edges):
        r, _ = rmat_edges(scale, b, min(total, b + chunk_edges), seed, device)
        deg += torch.bincount(r, minlength=n)
        del r
    return deg


def rmat_csr(scale: int, edge_factor: int, seed: int, dtype=torch.float32, device="cpu",
             off_dtype=torch.int32, chunk_edges: int = 1 << 26, row_begin: int = 0,
             row_end: Optional[int] = None):
    """R-MAT (0.57, 0.19, 0.19, 0.05), n = 2^scale rows, edge_factor * n edges, duplicates
    kept, entries of a row in edge order (stable sort by row).  With row_begin/row_end only
    the edges falling in that row block are kept (row-block sharding: every rank scans all
    edges in chunks and keeps its own)."""
    n = 1 << scale
    row_end = n if row_end is None else row_end
    total = edge_factor * n
    rows_l, cols_l, ids_l = [], [], []
    for b in range(0, total, chunk_edges):
        r, c = rmat_edges(scale, b, min(total, b + chunk_edges), seed, device)
        keep = (r >= row_begin) & (r < row_end)
        ids = torch.arange(b, min(total, b + chunk_edges), dtype=torch.int64, device=device)
        rows_l.append((r[keep] - row_begin).to(torch.int32))
        cols_l.append(c[keep].to(torch.int32))
        ids_l.append(ids[keep])
        del r, c, keep, ids
    rows_t = torch.cat(rows_l)
    cols_t = torch.cat(cols_l)
    ids_t = torch.cat(ids_l)
    del rows_l, cols_l, ids_l
    order = torch.sort(rows_t.to(torch.int64), stable=True).indices
    rows_s = rows_t[order]
    colind = cols_t[order].contiguous()
    ids_s = ids_t[order]
    del rows_t, cols_t, ids_t, order
    nrows = row_end - row_begin
    counts = torch.bincount(rows_s.to(torch.int64), minlength=nrows)
    rowptr = torch.zeros(nrows + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    if dtype == torch.int32:
        values = (_lsr(splitmix64(ids_s), 60) - 8).to(torch.int32)
    else:
        values = hash_uniform(seed + 7, ids_s, dtype)
    return values, rowptr.to(off_dtype), colind, (nrows, n)


def rmat_csr_blocked(scale: int, edge_factor: int, seed: int, deg: torch.Tensor,
                     dtype=torch.float32, device="cpu", off_dtype=torch.int32,
                     row_begin: int = 0, row_end: Optional[int] = None,
                     max_block_nnz: int = 1 << 29, chunk_edges: int = 1 << 26):
    """rmat_csr(...) of rows [row_begin, row_end) assembled from consecutive row sub-blocks of
    at most ~max_block_nnz entries each (cut with the row degrees `deg` of the whole graph,
    rmat_degrees): the same matrix, entry for entry, with the sort's temporaries bounded by
    the sub-block instead of the block — scale 27 on ONE GPU is 2^31 entries, and sorting
    them at once would need > 100 GB of temporaries.  Every sub-block scans all edges."""
    n = 1 << scale
    row_end = n if row_end is None else row_end
    rp_all = torch.zeros(row_end - row_begin + 1, dtype=torch.int64, device=deg.device)
    torch.cumsum(deg[row_begin:row_end], 0, out=rp_all[1:])
    nnz = int(rp_all[-1])
    pieces = max(1, -(-nnz // max_block_nnz))
    if pieces == 1:
        return rmat_csr(scale, edge_factor, seed, dtype, device, off_dtype, chunk_edges,
                        row_begin, row_end)
    targets = torch.tensor([(p * nnz) // pieces for p in range(1, pieces)], dtype=torch.int64,
                           device=rp_all.device)
    cuts = [0] + torch.searchsorted(rp_all, targets).tolist() + [row_end - row_begin]
    vals, cols = [], []
    for p in range(pieces):
        a, b = row_begin + cuts[p], row_begin + max(cuts[p + 1], cuts[p])
        if b <= a:
            continue
        v, _, c, _ = rmat_csr(scale, edge_factor, seed, dtype, device, torch.int64, chunk_edges, a, b)
        vals.append(v)
        cols.append(c)
    values = torch.cat(vals) if vals else torch.empty(0, dtype=dtype, device=device)
    del vals
    colind = torch.cat(cols) if cols else torch.empty(0, dtype=torch.int32, device=device)
    del cols
    assert int(values.numel()) == nnz
    return values, rp_all.to(off_dtype).to(device), colind, (row_end - row_begin, n)


def to_csc(values, rowptr, colind, shape):
    """Column-major image of a CSR matrix (for CSC tests): stable sort by column, so a
    column's entries are in ascending row order."""
    m, n = shape
    rows = torch.repeat_interleave(torch.arange(m, device=values.device),
                                   (rowptr[1:] - rowptr[:-1]).to(torch.int64))
    order = torch.sort(colind.to(torch.int64), stable=True).indices
    counts = torch.bincount(colind.to(torch.int64), minlength=n)
    colptr = torch.zeros(n + 1, dtype=torch.int64, device=values.device)
    torch.cumsum(counts, 0, out=colptr[1:])
    return (values[order].contiguous(), colptr.to(rowptr.dtype),
            rows[order].to(colind.dtype).contiguous(), (m, n))
