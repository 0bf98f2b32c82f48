"""The other BASELINE.json configs beside bench.py's headline (C2), as blocks of the same JSON
line: C1, C3 (k = 32, 128), C4 on one GPU and C5 (R-MAT scale 27, SpMV and SpMM) STRONG-scaled
over the ranks of the run.  Every block carries what the headline carries — device time,
GFLOP/s, algorithmic GB/s and its fraction of the HBM peak (SURVEY §8d formulae), the
reference's CPU multiply timed on the host (`cpu_baseline`, kind "reference" = oracle/_ref, the
real spblas::multiply compiled from /root/reference), the end-to-end figure through the
public API with host buffers (`e2e`), and `parity`: the device result of the timed-shape
product compared with the reference's result on the same operands under the north-star bound
|dy_i| <= (len_i + 2) eps sum_j |alpha a_ij x_j|  (max_err_over_tol <= 1 passes).

The oracle is used here only as the checker and as the CPU baseline; every measured product
goes through the public host API -> C ABI -> sm_100a kernels.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

EPS = {torch.float32: 2.0 ** -23, torch.float64: 2.0 ** -52}


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def time_loop(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def spmv_bytes(nnz, m, n, sT, sI, sO):
    return nnz * (sT + sI) + (m + 1) * sO + n * sT + m * sT


def spmm_bytes(nnz, m, n, k, sT, sI, sO):
    return nnz * (sT + sI) + (m + 1) * sO + n * k * sT + m * k * sT


def _oracle():
    from oracle import oracle as O
    O.build()
    return O, ("reference" if O.have_ref() else "oracle"), ("reference" if O.have_ref() else "port")


def _host_rows(rp, ci, v, r0, r1):
    """rows [r0, r1) of a device CSR matrix as host arrays with rowptr rebased to 0."""
    rph = rp[r0:r1 + 1].cpu().numpy()
    b, e = int(rph[0]), int(rph[-1])
    # (offsets are absolute into colind/values only when rowptr[0] == 0 — true for every
    # matrix built here)
    return (rph - rph[0]), ci[b:e].cpu().numpy(), v[b:e].cpu().numpy()


def _ref_call(O, fn, impl, *args, **kw):
    try:
        return fn(*args, impl=impl, **kw), impl
    except AttributeError:              # a type combination the reference shim does not export
        return fn(*args, impl="oracle", **kw), "oracle"


def check_spmv(rp, ci, v, x, alpha, y_dev, rows, n, what, reps=1):
    """The reference's CPU multiply on rows [r0, r1) against the same x: returns
    (cpu_baseline block, parity block).  `y_dev`: the device result for those rows."""
    O, impl, kind = _oracle()
    r0, r1 = rows
    rph, cih, vh = _host_rows(rp, ci, v, r0, r1)
    xh = x.cpu().numpy()
    best, yref = None, None
    for _ in range(max(1, reps)):
        t0 = time.perf_counter()
        yref, used = _ref_call(O, O.spmv, impl, "csr", (r1 - r0, n), rph, cih, vh, xh, alpha_a=alpha)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    kind = "reference" if used == "reference" else "port"
    nnz_r = int(rph[-1])
    bound = O.abs_rowsum(rph, cih, vh, xh, 1.0 if alpha is None else float(alpha))
    tol = (np.diff(rph).astype(np.float64) + 2.0) * EPS[v.dtype] * bound
    got = y_dev.cpu().numpy().astype(np.float64)
    err = np.abs(got - yref.astype(np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(tol > 0, err / tol, np.where(err > 0, np.inf, 0.0))
    finite = bool(np.isfinite(got).all())
    cpu = {"value": 2.0 * nnz_r / best / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
           "seconds": best,
           "sample": f"{what}: rows [{r0}, {r1}) ({nnz_r} stored entries), best of {max(1, reps)}; the "
                     "reference's CPU multiply is serial (1 thread)",
           "host_cores_available": os.cpu_count()}
    parity = {"max_err_over_tol": float(ratio.max()) if ratio.size else 0.0,
              "rows_checked": int(r1 - r0), "impl": kind, "finite": finite,
              "bound": "(len_i + 2) * eps * sum_j |alpha a_ij x_j|",
              "pass": bool(finite and (ratio.size == 0 or ratio.max() <= 1.0))}
    return cpu, parity


def check_spmm(rp, ci, v, B, alpha, C_dev, rows, n, what, compact=False):
    """compact: B is too large for the host (C5: 34 GB) — the reference multiplies the sampled
    rows against the rows of B they reference, renumbered in ascending order: the same products
    in the same order."""
    O, impl, kind = _oracle()
    r0, r1 = rows
    rph, cih, vh = _host_rows(rp, ci, v, r0, r1)
    if compact:
        cols = np.unique(cih)
        cih = np.searchsorted(cols, cih).astype(cih.dtype)
        Bh = B[torch.from_numpy(cols.astype(np.int64)).to(B.device)].cpu().numpy()
        n = len(cols)
        what += f" (against the {n} rows of B it references)"
    else:
        Bh = B.cpu().numpy()
    k = Bh.shape[1]
    t0 = time.perf_counter()
    Cref, used = _ref_call(O, O.spmm, impl, "csr", (r1 - r0, n), rph, cih, vh, Bh, alpha_a=alpha)
    sec = time.perf_counter() - t0
    kind = "reference" if used == "reference" else "port"
    nnz_r = int(rph[-1])
    a = 1.0 if alpha is None else float(alpha)
    if a > 0 and (vh >= 0).all() and (Bh >= 0).all():
        bound = np.abs(Cref.astype(np.float64))       # every term is non-negative: the sum is its own bound
    else:
        bound = np.abs(O.spmm("csr", (r1 - r0, n), rph, cih, np.abs(vh), np.abs(Bh),
                              alpha_a=abs(a) if alpha is not None else None).astype(np.float64))
    # (the bound itself carries a relative rounding error of len * eps: one more unit covers it)
    tol = (np.diff(rph).astype(np.float64)[:, None] + 3.0) * EPS[v.dtype] * bound
    got = C_dev.cpu().numpy().astype(np.float64)
    err = np.abs(got - Cref.astype(np.float64))
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(tol > 0, err / tol, np.where(err > 0, np.inf, 0.0))
    finite = bool(np.isfinite(got).all())
    cpu = {"value": 2.0 * nnz_r * k / sec / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
           "seconds": sec,
           "sample": f"{what}: rows [{r0}, {r1}) ({nnz_r} stored entries) x k={k}, once; the "
                     "reference's CPU multiply is serial (1 thread)",
           "host_cores_available": os.cpu_count()}
    parity = {"max_err_over_tol": float(ratio.max()) if ratio.size else 0.0,
              "rows_checked": int(r1 - r0), "impl": kind, "finite": finite,
              "bound": "(len_i + 3) * eps * sum_k |alpha a_ik B_kj|",
              "pass": bool(finite and (ratio.size == 0 or ratio.max() <= 1.0))}
    return cpu, parity


def e2e_spmv(sb, info, a, x_dev, m, flops, steps=5):
    """End to end through the host-buffer entry point: pinned x in, pinned y out, every step."""
    x_host = x_dev.cpu().pin_memory()
    y_host = torch.empty(m, dtype=x_dev.dtype).pin_memory()

    def step():
        sb.multiply_execute_host(info, a, x_host, y_host)
        torch.cuda.current_stream().synchronize()
    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = (time.perf_counter() - t0) / steps * 1e3
    return {"value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms, "steps": steps,
            "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size()),
            "d2h_bytes_per_step": int(y_host.numel() * y_host.element_size()),
            "what": "multiply_execute_host (spblas_b200_spmv_host): pinned host x -> device, "
                    "kernels, y -> pinned host, pipelined over chunks of the partition; A and "
                    "the plan stay resident"}, y_host


def e2e_spmm(sb, info, a, B_dev, C_dev, flops, steps=3):
    """SpMM has no host-buffer entry point: copy B in, multiply_execute, copy C out."""
    B_host = B_dev.cpu().pin_memory()
    C_host = torch.empty(C_dev.shape, dtype=C_dev.dtype).pin_memory()
    Bd = torch.empty_like(B_dev)

    def step():
        Bd.copy_(B_host, non_blocking=True)
        sb.multiply_execute(info, a, Bd, C_dev)
        C_host.copy_(C_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = (time.perf_counter() - t0) / steps * 1e3
    return {"value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms, "steps": steps,
            "h2d_bytes_per_step": int(B_host.numel() * B_host.element_size()),
            "d2h_bytes_per_step": int(C_host.numel() * C_host.element_size()),
            "what": "pinned host B -> device, multiply_execute, C -> pinned host, one after the "
                    "other on one stream; A and the plan stay resident"}


def roofline(nbytes, ms, peak, peak_src, traffic, kernel):
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "kernel": kernel, "algorithmic_bytes_per_launch": nbytes,
            "traffic_over_algorithmic": (traffic / nbytes) if traffic else None,
            "kernel_ms": ms, "peak_source": peak_src}


SPMV_KERNEL = {0: "spmv_merge_tile_kernel", 1: "spmv_pipe_kernel", 2: "spmv_warp_stream_kernel",
               3: "spmv_hub_stream_kernel", 4: "spmv_hubg_stream_kernel"}


def _progress(ctx, msg):
    """stderr breadcrumbs of the long blocks (SPBLAS_B200_BENCH_VERBOSE=1)."""
    if os.environ.get("SPBLAS_B200_BENCH_VERBOSE") == "1" and ctx.get("rank", 0) == 0:
        import sys
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# single-GPU configs
# --------------------------------------------------------------------------------------------
def config_c1(ctx):
    """C1: examples/simple_spmv-style uniform random CSR SpMV fp32/int32, m = n = 1M, 10 per row,
    scaled(1.2, a) as the example does; cold-L2 protocol (92 MB of operands fit in L2: the loop
    rotates over 4 independent operand sets); timed through multiply_execute(info, ...) AND
    through the no-info multiply(a, x, y) the reference's example and device test use."""
    sb, G, dev, K, W = ctx["sb"], ctx["G"], ctx["dev"], ctx["K"], ctx["W"]
    m = n = 1_000_000
    copies = 4
    mats = []
    for c in range(copies):
        v, rp, ci, shape = G.uniform_random_csr(m, n, 10, seed=c, dtype=torch.float32, device=dev)
        a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
        x = G.dense_uniform((n,), 100 + c, torch.float32, dev)
        y = torch.empty(m, device=dev)
        mats.append((a, x, y, sb.multiply_inspect(a, x, y)))
    nnz = mats[0][0].nnz

    def fn(i):
        a, x, y, info = mats[i % copies]
        sb.multiply_execute(info, sb.scaled(1.2, a), x, y)

    def fn_noinfo(i):
        a, x, y, _ = mats[i % copies]
        sb.multiply(sb.scaled(1.2, a), x, y)
    ms = time_loop(fn, K, W)
    a, x, y, info = mats[0]
    fn(0)
    torch.cuda.synchronize()
    y_info = y.clone()
    ms_noinfo = time_loop(fn_noinfo, K, W)
    fn_noinfo(0)
    torch.cuda.synchronize()
    same = bool(torch.equal(y, y_info))
    cpu, parity = check_spmv(a.rowptr, a.colind, a.values, x, 1.2, y_info, (0, m), n, "C1 full product", reps=3)
    parity["no_info_overload_bit_identical"] = same
    flops, nbytes = 2.0 * nnz, spmv_bytes(nnz, m, n, 4, 4, 4)
    e2e, _ = e2e_spmv(sb, info, sb.scaled(1.2, a), x, m, flops)
    variant = info.spmv_variant
    blk = {"workload": "C1 uniform random CSR SpMV fp32/int32 m=n=1M, 10 nnz/row, scaled(1.2, a)",
           "nnz": nnz, "ms": ms, "gflops": flops / ms / 1e6, "gbs": nbytes / ms / 1e6,
           "l2_policy": f"rotating over {copies} independent operand sets (cold L2)",
           "no_info_overload": {"ms": ms_noinfo, "gflops": flops / ms_noinfo / 1e6,
                                "what": "multiply(a, x, y) — the spelling of examples/simple_spmv.cpp "
                                        "and test/gtest/device/spmv_test.cpp:34: partition derived on "
                                        "every call", "overhead_us": (ms_noinfo - ms) * 1e3},
           "roofline": roofline(nbytes, ms, ctx["peak"], ctx["peak_src"], ctx["traffic"]("c1"),
                                SPMV_KERNEL.get(variant, str(variant))),
           "cpu_baseline": cpu, "e2e": e2e, "parity": parity}
    for t in mats:
        t[3].close()
    return blk


def config_c3(ctx, k):
    """C3: CSR SpMM fp32, random 2M x 2M at 16 nnz/row times dense row-major B, k = 32 / 128."""
    sb, G, dev, K, W = ctx["sb"], ctx["G"], ctx["dev"], ctx["K"], ctx["W"]
    m = n = 2_000_000
    v, rp, ci, shape = G.uniform_random_csr(m, n, 16, seed=3, dtype=torch.float32, device=dev)
    nnz = int(ci.numel())
    a = sb.csr_view(v, rp, ci, shape, nnz)
    B = G.dense_uniform((n, k), 4, torch.float32, dev)
    C = torch.empty((m, k), device=dev)
    info = sb.multiply_inspect(a, B, C)
    ms = time_loop(lambda i: sb.multiply_execute(info, a, B, C), min(K, 20), W)
    flops, nbytes = 2.0 * nnz * k, spmm_bytes(nnz, m, n, k, 4, 4, 4)
    rows = (0, m) if k <= 32 else (0, 100_000)
    what = f"C3 k={k} " + ("full product" if rows[1] == m else "row sample")
    cpu, parity = check_spmm(rp, ci, v, B, None, C[rows[0]:rows[1]], rows, n, what)
    e2e = e2e_spmm(sb, info, a, B, C, flops)
    blk = {"workload": f"C3 CSR SpMM fp32 2M x 2M, 16 nnz/row, row-major B k={k}", "nnz": nnz,
           "ms": ms, "gflops": flops / ms / 1e6, "gbs": nbytes / ms / 1e6,
           "l2_policy": "inputs larger than L2",
           "spmm_variant": info.spmm_variant,
           "gather_model_bytes": nnz * 8 + (m + 1) * 4 + nnz * k * 4 + m * k * 4,
           "roofline": roofline(nbytes, ms, ctx["peak"], ctx["peak_src"], ctx["traffic"](f"c3k{k}"),
                                "spmm_row_kernel" if info.spmm_variant < 1000 else "spmm_ring_kernel"),
           "cpu_baseline": cpu, "e2e": e2e, "parity": parity}
    info.close()
    return blk


def config_c4(ctx):
    """C4: R-MAT scale 24 (edge factor 16) CSR SpMV fp32/int32 — skewed rows and columns."""
    sb, G, dev, K, W = ctx["sb"], ctx["G"], ctx["dev"], ctx["K"], ctx["W"]
    v, rp, ci, shape = G.rmat_csr(24, 16, seed=24, dtype=torch.float32, device=dev)
    m, n = shape
    nnz = int(ci.numel())
    a_plain = sb.csr_view(v, rp, ci, shape, nnz)
    a = sb.matrix_opt(a_plain)             # the reference's marker: the plan may keep optimised state
    x = G.dense_uniform((n,), 5, torch.float32, dev)
    y = torch.empty(m, device=dev)
    t0 = time.perf_counter()
    info = sb.multiply_inspect(a, x, y)
    sb.multiply_execute(info, a, x, y)      # first product builds the lazy tables
    torch.cuda.synchronize()
    setup_ms = (time.perf_counter() - t0) * 1e3
    ms = time_loop(lambda i: sb.multiply_execute(info, a, x, y), K, W)
    y2 = torch.empty_like(y)
    info2 = sb.multiply_inspect(a_plain, x, y2)
    ms_plain = time_loop(lambda i: sb.multiply_execute(info2, a_plain, x, y2), K, W)
    flops, nbytes = 2.0 * nnz, spmv_bytes(nnz, m, n, 4, 4, 4)
    cpu, parity = check_spmv(rp, ci, v, x, None, y, (0, m), n, "C4 full product")
    parity["plain_and_matrix_opt_bit_identical"] = bool(torch.equal(y, y2))
    e2e, _ = e2e_spmv(sb, info2, a_plain, x, m, flops)
    variant = info.spmv_variant
    blk = {"workload": "C4 R-MAT scale 24 (edge factor 16) CSR SpMV fp32/int32, operand wrapped in "
                       "matrix_opt", "nnz": nnz, "ms": ms, "gflops": flops / ms / 1e6,
           "gbs": nbytes / ms / 1e6, "l2_policy": "inputs larger than L2 (2.3 GB)",
           "inspect_plus_first_execute_ms": setup_ms, "max_row_len": info.max_row_len,
           "empty_rows": info.empty_rows, "spmv_variant": variant,
           "hub_columns": info.hub_count, "hub_reference_share": info.hub_refs / max(nnz, 1),
           "plain_operand": {"ms": ms_plain, "spmv_variant": info2.spmv_variant,
                             "gflops": flops / ms_plain / 1e6},
           "roofline": roofline(nbytes, ms, ctx["peak"], ctx["peak_src"],
                                ctx["traffic"]("c4" if variant == 3 else "c4_plain_walk"),
                                SPMV_KERNEL.get(variant, str(variant))),
           "cpu_baseline": cpu, "e2e": e2e, "parity": parity}
    info.close()
    info2.close()
    return blk


# --------------------------------------------------------------------------------------------
# C5: R-MAT scale 27, strong-scaled over the ranks
# --------------------------------------------------------------------------------------------
def _c5_block(ctx, scale):
    """This rank's nnz-balanced row block of the scale-`scale` R-MAT (fp64, int32 indices, int64
    offsets)."""
    G, dev, world, rank = ctx["G"], ctx["dev"], ctx["world"], ctx["rank"]
    from spblas_reference_b200.sharded import balanced_nnz_blocks
    n = 1 << scale
    deg = G.rmat_degrees(scale, 16, seed=27, device=dev)
    rowptr_all = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=rowptr_all[1:])
    blocks = balanced_nnz_blocks(rowptr_all, world)
    max_deg = int(deg.max())
    del rowptr_all
    r0, r1 = blocks[rank]
    v, rp, ci, shape = G.rmat_csr_blocked(scale, 16, 27, deg, dtype=torch.float64, device=dev,
                                          off_dtype=torch.int64, row_begin=r0, row_end=r1)
    del deg
    return v, rp, ci, shape, blocks, max_deg


def config_c5(ctx, scale=27, with_spmm=True):
    """C5: R-MAT scale 27 (2^31 stored entries, 134 M rows) CSR SpMV fp64 with int32 indices and
    int64 offsets, nnz-balanced row blocks over the ranks of the run (STRONG scaling: the matrix
    is fixed), x replicated, iterated y -> x with an allgather of the blocks; and the SpMM half
    (k = 32, B replicated, single product, no exchange).  Returns {"c5": ..., "c5mm": ...}."""
    import torch.distributed as dist
    from spblas_reference_b200.sharded import ShardedSpMV
    sb, G, dev, K, W = ctx["sb"], ctx["G"], ctx["dev"], ctx["K"], ctx["W"]
    world, rank = ctx["world"], ctx["rank"]
    barrier, max_over_ranks, sum_over_ranks = ctx["barrier"], ctx["max"], ctx["sum"]
    n = 1 << scale
    t0 = time.perf_counter()
    _progress(ctx, f"c5: generating scale {scale}")
    v, rp, ci, shape, blocks, max_deg = _c5_block(ctx, scale)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    _progress(ctx, f"c5: generated {int(ci.numel())} entries in {gen_s:.1f} s")
    m_loc, nnz_loc = shape[0], int(ci.numel())
    r0, r1 = blocks[rank]
    a_plain = sb.csr_view(v, rp, ci, shape, nnz_loc)
    # matrix_opt: the reference's marker for "the backend may keep optimised state for this
    # matrix" — here x at the most referenced columns in a compact table (x is 1.07 GB at scale
    # 27: a gather that misses L2 costs a DRAM sector)
    a = sb.matrix_opt(a_plain) if os.environ.get("SPBLAS_B200_MATRIX_OPT", "1") != "0" else a_plain
    alpha = 1.0 / max_deg                       # keeps the iterates in [0, 1]
    a_scaled = sb.scaled(alpha, a)
    x0 = G.dense_uniform((n,), 5, torch.float64, dev)
    t0 = time.perf_counter()
    y_tmp = torch.empty(m_loc, dtype=torch.float64, device=dev)
    info = sb.multiply_inspect(a, x0, y_tmp)
    torch.cuda.synchronize()
    inspect_ms = (time.perf_counter() - t0) * 1e3
    _progress(ctx, f"c5: inspected in {inspect_ms:.1f} ms")
    t0 = time.perf_counter()
    sb.multiply_execute(info, a_scaled, x0, y_tmp)      # first product: builds the lazy tables
    torch.cuda.synchronize()
    first_ms = (time.perf_counter() - t0) * 1e3
    _progress(ctx, f"c5: first product {first_ms:.1f} ms, variant {info.spmv_variant}, hubs {info.hub_count}")
    # the plain operand (no matrix_opt) beside it: kernel only
    info_plain = sb.multiply_inspect(a_plain, x0, y_tmp)
    plain_ms = max_over_ranks(time_loop(
        lambda i: sb.multiply_execute(info_plain, sb.scaled(alpha, a_plain), x0, y_tmp), 5, 2))
    plain_variant = info_plain.spmv_variant
    info_plain.close()
    del y_tmp
    op = ShardedSpMV(n, blocks, (0, n), lambda x, y: sb.multiply_execute(info, a_scaled, x, y),
                     torch.float64, dev, info=info,
                     fused=None if os.environ.get("SPBLAS_B200_FUSED", "1") != "0" else False,
                     multicast={"1": True, "0": False}.get(os.environ.get("SPBLAS_B200_MULTICAST")))
    _progress(ctx, f"c5: plain operand {plain_ms:.3f} ms; exchange {op.exchange_impl}")
    op.set_x(x0)
    del x0
    Kc = max(3, min(K, 10))
    for _ in range(max(3, min(W, 5))):
        op.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    for _ in range(2 if world > 1 else 0):      # untimed: the fused barrier aligns the GPUs (bench.py)
        op.step()
    l0 = info.total_launches
    e0.record()
    for _ in range(Kc):
        op.step()
    e1.record()
    barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1) / Kc)
    launches = int(sum_over_ranks(info.total_launches - l0))
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    k0.record()
    for _ in range(Kc):
        op.multiply()
    k1.record()
    barrier()
    kern_ms = max_over_ranks(k0.elapsed_time(k1) / Kc)

    variant_used = info.spmv_variant          # (the host-buffer e2e leg below runs the plain walk)
    hub_columns, hub_refs = info.hub_count, info.hub_refs
    _progress(ctx, f"c5: step {step_ms:.3f} ms, kernels {kern_ms:.3f} ms")
    # ---- parity after >= 3 fused iterations: (1) this rank's replica of x is bit for bit what an
    # out-of-band allgather of the blocks gives, (2) its rows of the next product are within the
    # bound of the reference's CPU multiply fed with the same x (per iteration, not compounded)
    x_in = op.x_current.clone()
    replica_ok = True
    if world > 1:
        mine = x_in[r0:r1].contiguous()
        sizes = [e - b for b, e in blocks]
        parts = [torch.empty(s, dtype=torch.float64, device=dev) for s in sizes]
        # (allgather of uneven blocks: one broadcast per block)
        for src, (b, e) in enumerate(blocks):
            buf = mine if src == rank else parts[src]
            if e > b:
                dist.broadcast(buf, src=src)
            if src != rank and e > b:
                replica_ok = replica_ok and bool(torch.equal(buf, x_in[b:e]))
        del parts
    y_blk = op.step().clone()
    rows_n = int(min(m_loc, max(1, 200_000)))
    # the sample: the block's first rows (rank 0's are the densest of the matrix)
    cpu, parity = check_spmv(rp, ci, v, x_in, alpha, y_blk[:rows_n], (0, rows_n), n,
                             f"C5 scale {scale} rank {rank} row sample")
    # ... and its LAST rows: at N = 1 their offsets lie beyond 2^31 - 1 (the reason this config
    # has 64-bit offsets)
    if m_loc > rows_n:
        t0_, t1_ = max(rows_n, m_loc - rows_n), m_loc
        _, par_tail = check_spmv(rp, ci, v, x_in, alpha, y_blk[t0_:t1_], (t0_, t1_), n, "tail rows")
        parity["max_err_over_tol"] = max(parity["max_err_over_tol"], par_tail["max_err_over_tol"])
        parity["rows_checked"] += par_tail["rows_checked"]
        parity["pass"] = parity["pass"] and par_tail["pass"]
        parity["largest_offset_checked"] = int(rp[-1])
    parity["max_err_over_tol"] = max_over_ranks(parity["max_err_over_tol"])
    parity["rows_checked"] = int(sum_over_ranks(parity["rows_checked"]))
    parity["pass"] = bool(max_over_ranks(0.0 if parity["pass"] else 1.0) == 0.0)
    parity["replica_bit_identical_to_allgather"] = bool(max_over_ranks(0.0 if replica_ok else 1.0) == 0.0)
    parity["pass"] = parity["pass"] and parity["replica_bit_identical_to_allgather"]
    parity["what"] = ("every rank: its replica of x after the fused iterations == an out-of-band "
                      "allgather of the blocks (bit for bit), and the first rows of its block of the "
                      "next product against the reference's CPU multiply on the same x")
    timeout_flag = int(max_over_ranks(float(info.barrier_timeout))) if op.fused else 0

    total_nnz = int(sum_over_ranks(nnz_loc))
    nbytes = spmv_bytes(nnz_loc, m_loc, n, 8, 4, 8)      # full replicated x
    flops = 2.0 * total_nnz
    if op.fused:                                          # plain products again
        info.set_scatter(())
        info.set_barrier((), ())
    _progress(ctx, f"c5: parity {parity['max_err_over_tol']:.3f} pass {parity['pass']}")
    # ---- end to end: every rank uploads x and downloads its block, every step
    e2e, _ = e2e_spmv(sb, info, a_scaled, x_in, m_loc, flops, steps=3)
    e2e["ms_per_step"] = max_over_ranks(e2e["ms_per_step"])
    e2e["value"] = flops / (e2e["ms_per_step"] * 1e-3) / 1e9
    e2e["h2d_bytes_per_step"] = int(sum_over_ranks(e2e["h2d_bytes_per_step"]))
    e2e["d2h_bytes_per_step"] = int(sum_over_ranks(e2e["d2h_bytes_per_step"]))
    del x_in, y_blk
    exchange_bytes = (n - (r1 - r0)) * 8
    c5 = {"workload": f"C5 R-MAT scale {scale} (edge factor 16) CSR SpMV fp64, int32 indices, int64 "
                      f"offsets, {world} nnz-balanced row block(s), iterated y->x",
          "scaling": "strong", "n_gpus": world, "nnz": total_nnz, "rows": n,
          "rows_rank0": m_loc if rank == 0 else None, "nnz_rank0": nnz_loc if rank == 0 else None,
          "ms": step_ms, "gflops": flops / step_ms / 1e6, "steps": Kc,
          "kernel_only_ms": kern_ms, "kernel_only_gflops": flops / kern_ms / 1e6,
          "generate_s": gen_s, "inspect_ms": inspect_ms, "first_execute_ms": first_ms,
          "max_row_len": info.max_row_len,
          "spmv_variant": variant_used, "gpu_launches": launches,
          "hub_columns": hub_columns, "hub_reference_share": hub_refs / max(nnz_loc, 1),
          "plain_operand": {"kernel_only_ms": plain_ms, "spmv_variant": plain_variant,
                            "what": "the same product without matrix_opt (warp-stream kernel)"},
          "exchange": {"mode": op.plan.mode,
                       "impl": op.exchange_impl, "calibration": op.calibration,
                       "bytes_received_per_gpu_per_step": exchange_bytes if world > 1 else 0,
                       "fused_error": op.fused_error, "barrier_timeout_flag": timeout_flag,
                       "ms_above_kernels": step_ms - kern_ms},
          "l2_policy": "inputs larger than L2",
          "roofline": roofline(nbytes, kern_ms, ctx["peak"], ctx["peak_src"],
                               ctx["traffic"](f"c5_n{world}"),
                               SPMV_KERNEL.get(variant_used, "?") + "<double,int,long>"
                               + (" (+ hub_fill_kernel)" if variant_used == 4 else "")),
          "cpu_baseline": cpu if rank == 0 else None, "e2e": e2e, "parity": parity}
    c5["roofline"]["note"] = ("per GPU: this rank-0-timed launch streams its block of A and gathers "
                              "from the full replicated x (1.07 GB at scale 27: beyond L2, so a gather "
                              "that misses costs a 32-byte DRAM sector — SURVEY 8d caveat)")
    out = {"c5": c5}
    del op
    if not with_spmm:
        info.close()
        return out

    # ---- the SpMM half: C = A B, k = 32, B replicated, no exchange -------------------------
    _progress(ctx, f"c5: e2e {e2e['ms_per_step']:.1f} ms; SpMM half next")
    k = 32
    torch.cuda.empty_cache()
    B = G.dense_uniform_rows(n, k, 6, torch.float64, dev)
    C = torch.empty((m_loc, k), dtype=torch.float64, device=dev)
    a = a_plain
    info_mm = sb.multiply_inspect(a, B, C)
    Km = max(2, min(K, 4))
    for i in range(2):
        sb.multiply_execute(info_mm, a, B, C)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(Km):
        sb.multiply_execute(info_mm, a, B, C)
    e1.record()
    barrier()
    mm_ms = max_over_ranks(e0.elapsed_time(e1) / Km)
    _progress(ctx, f"c5mm: {mm_ms:.2f} ms, variant {info_mm.spmm_variant}")
    rows_mm = int(min(m_loc, 20_000))
    cpu_mm, par_mm = check_spmm(rp, ci, v, B, None, C[:rows_mm], (0, rows_mm), n,
                                f"C5 SpMM scale {scale} rank {rank} row sample", compact=True)
    par_mm["max_err_over_tol"] = max_over_ranks(par_mm["max_err_over_tol"])
    par_mm["rows_checked"] = int(sum_over_ranks(par_mm["rows_checked"]))
    par_mm["pass"] = bool(max_over_ranks(0.0 if par_mm["pass"] else 1.0) == 0.0)
    mm_flops = 2.0 * total_nnz * k
    mm_bytes = spmm_bytes(nnz_loc, m_loc, n, k, 8, 4, 8)
    # end to end on a bounded part: B is 34 GB per GPU — the host leg moves this rank's C block only
    C_host = torch.empty((min(m_loc, 1 << 20), k), dtype=torch.float64).pin_memory()
    t0 = time.perf_counter()
    sb.multiply_execute(info_mm, a, B, C)
    C_host.copy_(C[:C_host.shape[0]], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    mm_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    out["c5mm"] = {
        "workload": f"C5 R-MAT scale {scale} CSR SpMM fp64, k={k}, int32 indices, int64 offsets, "
                    f"{world} nnz-balanced row block(s), B replicated, single product (no exchange)",
        "scaling": "strong", "n_gpus": world, "nnz": total_nnz, "ms": mm_ms, "steps": Km,
        "gflops": mm_flops / mm_ms / 1e6, "spmm_variant": info_mm.spmm_variant,
        "num_segments": info_mm.num_segments,
        "exchange": {"mode": "none", "bytes_received_per_gpu_per_step": 0},
        "roofline": roofline(mm_bytes, mm_ms, ctx["peak"], ctx["peak_src"],
                             ctx["traffic"](f"c5mm_n{world}"), "spmm (see spmm_variant)"),
        "cpu_baseline": cpu_mm if rank == 0 else None,
        "e2e": {"value": mm_flops / (mm_e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
                "ms_per_step": mm_e2e_ms, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": int(C_host.numel() * 8) * world,
                "what": "B (34 GB per GPU, replicated) stays resident; the product plus the "
                        "download of the first 2^20 rows of every rank's C block to pinned host "
                        "memory"},
        "parity": par_mm}
    info_mm.close()
    info.close()
    return out
