"""Side workloads for bench.py (--workload c1|c3k32|c3k128|c4): the other BASELINE.json
configs on ONE GPU.  They are parity-test cases, not the headline bench line; this prints
the same JSON shape so that profiles/ can hold their roofline numbers too.

C1 is 92 MB — it fits in the 126 MB L2 — so its timed loop rotates over enough independent
copies of the operands to exceed 2x L2 (cold-L2 protocol)."""
import json
import os
import sys
import time

import torch


def cusparse_compare(kind, operands, steps, scale=1.0, transpose=False):
    """ms per product of NVIDIA cuSPARSE — what the reference's NVIDIA backend calls
    (vendor/cusparse/spmv_impl.hpp:80-84, CUSPARSE_SPMV_ALG_DEFAULT) — on the same operands,
    same box, same timing loop (scripts/cusparse_comparator.py; comparator only).  `operands`
    is a list of (m, n, rowptr, colind, values, x_or_B, y_or_C) rotated like the timed loop."""
    if os.environ.get("SPBLAS_B200_NO_CUSPARSE", "0") == "1":
        return None
    try:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
        import cusparse_comparator as cc
        cs = cc.CuSparse()
        out = {}
        algs = {"spmv": (("CUSPARSE_SPMV_ALG_DEFAULT", 0), ("CUSPARSE_SPMV_CSR_ALG2", 3)),
                "spmm": (("CUSPARSE_SPMM_ALG_DEFAULT", 0), ("CUSPARSE_SPMM_CSR_ALG2", 6))}[kind]
        for name, alg in algs:
            runs = []
            for (m, n, rp, ci, v, xin, yout) in operands:
                if kind == "spmv":
                    runs.append(cs.spmv(m, n, rp, ci, v, xin, yout, alpha=scale, alg=alg,
                                        transpose=transpose))
                else:
                    runs.append(cs.spmm(m, n, xin.shape[1], rp, ci, v, xin, yout, alpha=scale, alg=alg))
            state = {"i": 0}

            def run():
                runs[state["i"] % len(runs)]()
                state["i"] += 1
            out[name] = cc.time_ms(run, steps)
        out["note"] = ("comparator only: workspace preallocated, only cusparseSpMV/SpMM timed; the "
                       "reference's wrapper also pays bufferSize + cudaMalloc + cudaFree per call")
        return out
    except Exception as e:                       # the comparator must never break the bench
        return {"unavailable": f"{type(e).__name__}: {e}"}


def _bytes_spmv(nnz, m, n, sT, sI=4, sO=4):
    return nnz * (sT + sI) + (m + 1) * sO + n * sT + m * sT


def _bytes_spmm(nnz, m, n, k, sT, sI=4, sO=4):
    return nnz * (sT + sI) + (m + 1) * sO + n * k * sT + m * k * sT


def cpu_baseline(kind, m, n, rp, ci, v, xin, alpha, rows_sample, flops_per_nnz, what):
    """The reference's own CPU multiply (oracle/_ref when built from /root/reference, else the
    oracle port) on a bounded sample — the first `rows_sample` rows of the same matrix against
    the same dense operand — on the box's host cores (1 thread: the reference is serial)."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    R = int(min(m, rows_sample))
    rph = rp[:R + 1].cpu().numpy()
    nnz_r = int(rph[-1] - rph[0])
    cih, vh = ci[:nnz_r + int(rph[0])].cpu().numpy(), v[:nnz_r + int(rph[0])].cpu().numpy()
    xh = xin.cpu().numpy()
    impl, kindname = ("reference", "reference") if O.have_ref() else ("oracle", "port")
    fn = O.spmv if kind == "spmv" else O.spmm
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        try:
            fn("csr", (R, n), rph, cih, vh, xh, alpha_a=alpha, impl=impl)
        except AttributeError:              # type combination the reference shim does not export
            impl, kindname = "oracle", "port"
            fn("csr", (R, n), rph, cih, vh, xh, alpha_a=alpha, impl=impl)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": flops_per_nnz * nnz_r / best / 1e9, "unit": "GFLOP/s", "cores": 1,
            "kind": kindname, "seconds": best,
            "sample": f"{what}: rows [0, {R}) of {m} ({nnz_r} stored entries), best of 2; the "
                      "reference CPU multiply is serial (1 thread)",
            "host_cores_available": os.cpu_count()}


def gather_ceiling(sb_cabi, operands, steps, with_values=True):
    """ms of the gather probe (csrc/probe.cu) on the workload's own colind / values / x: the
    same loads as SpMV and nothing else.  `operands`: list of (colind, values, x) rotated
    like the timed loop.  Best over the probe's occupancy settings."""
    import ctypes as C
    L = sb_cabi.lib()
    ci0, v0, x0 = operands[0]
    vt = sb_cabi.F32 if v0.dtype == torch.float32 else sb_cabi.F64
    it = sb_cabi.I32 if ci0.dtype == torch.int32 else sb_cabi.I64
    out = torch.empty(148 * 8 * 256 * 2, dtype=v0.dtype, device=v0.device)
    stream = torch.cuda.current_stream().cuda_stream
    best = None
    for ctas in (4, 6, 8):
        def fn(i):
            ci, v, x = operands[i % len(operands)]
            st = L.spblas_b200_probe_gather(stream, it, vt, ci.numel(), ci.data_ptr(),
                                            v.data_ptr() if with_values else None,
                                            x.data_ptr(), out.data_ptr(), ctas)
            assert st == 0, st
        ms = _time_loop(fn, steps, 3)
        best = ms if best is None else min(best, ms)
    return best


def _time_loop(fn, steps, warmup):
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


def run_extra(args, sb, G, dev, peak, peak_src, sampler):
    K, W = max(1, args.steps), max(3, args.warmup)
    wl = args.workload
    extra = {}
    if wl == "c1":
        m = n = 1_000_000
        copies = 4                                    # 4 x 92 MB > 2 x L2
        mats = []
        for c in range(copies):
            v, rp, ci, shape = G.uniform_random_csr(m, n, 10, seed=c, dtype=torch.float32, device=dev)
            a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
            x = torch.ones(n, device=dev)
            y = torch.empty(m, device=dev)
            mats.append((a, x, y, sb.multiply_inspect(a, x, y)))
        nnz = mats[0][0].nnz

        def fn(i):
            a, x, y, info = mats[i % copies]
            sb.multiply_execute(info, sb.scaled(1.2, a), x, y)
        flops, nbytes, dtype = 2.0 * nnz, _bytes_spmv(nnz, m, n, 4), "f32"
        name = "C1 uniform random CSR SpMV fp32/int32 m=n=1M, 10 nnz/row, scaled(1.2, a)"
        extra["l2_policy"] = f"rotating over {copies} independent operand sets (cold L2)"
        launches_of = lambda: sum(t[3].total_launches for t in mats)
        cmp_args = ("spmv", [(m, n, t[0].rowptr, t[0].colind, t[0].values, t[1],
                              torch.empty_like(t[2])) for t in mats], 1.2)
        result_of = lambda: mats[0][2]
        probe_ops = [(t[0].colind, t[0].values, t[1]) for t in mats]
        cpu_args = ("spmv", m, n, mats[0][0].rowptr, mats[0][0].colind, mats[0][0].values,
                    mats[0][1], 1.2, m, 2.0, "C1 full product")
    elif wl in ("c1t", "c4t"):
        # y = A^T x through transposed(a) (the second half of the alternating
        # A x / A^T y loop of notes/spmv.hpp:14-22): the inspect phase builds the row-major
        # image of A^T once, the execute kernels gather the values through its permutation
        if wl == "c1t":
            v, rp, ci, shape = G.uniform_random_csr(1_000_000, 1_000_000, 10, seed=0,
                                                    dtype=torch.float32, device=dev)
            name = "C1 matrix, y = A^T x via transposed(a), fp32/int32"
        else:
            v, rp, ci, shape = G.rmat_csr(22, 16, seed=24, dtype=torch.float32, device=dev)
            name = "R-MAT scale 22 (edge factor 16), y = A^T x via transposed(a), fp32/int32"
        m, n = shape
        nnz = int(ci.numel())
        a = sb.transposed(sb.csr_view(v, rp, ci, shape, nnz))          # n x m
        if os.environ.get("SPBLAS_B200_MATRIX_OPT", "1") != "0":
            # matrix_opt: the plan may keep the values gathered in image order (the solver
            # loop's values are static); =0 measures the gather-through-permutation path
            a = sb.matrix_opt(a)
            name += ", matrix_opt (values cached at inspect)"
        x = G.dense_uniform((m,), 5, torch.float32, dev)
        y = torch.empty(n, device=dev)
        t0 = time.perf_counter()
        info = sb.multiply_inspect(a, x, y)
        torch.cuda.synchronize()
        extra["inspect_ms"] = (time.perf_counter() - t0) * 1e3

        def fn(i):
            sb.multiply_execute(info, a, x, y)
        flops, nbytes, dtype = 2.0 * nnz, _bytes_spmv(nnz, n, m, 4), "f32"
        extra["l2_policy"] = "inputs larger than L2"
        launches_of = lambda: info.total_launches
        cmp_args = ("spmv", [(m, n, rp, ci, v, x, torch.empty_like(y))], 1.0)
        cmp_transpose = True
        result_of = lambda: y
        probe_ops = None
        cpu_args = None
    elif wl == "c4":
        v, rp, ci, shape = G.rmat_csr(24, 16, seed=24, dtype=torch.float32, device=dev)
        m, n = shape
        nnz = int(ci.numel())
        a_plain = sb.csr_view(v, rp, ci, shape, nnz)
        name = "C4 R-MAT scale 24 (edge factor 16) CSR SpMV fp32/int32"
        a = a_plain
        if os.environ.get("SPBLAS_B200_MATRIX_OPT", "1") != "0":
            # matrix_opt: the plan may keep structure-derived state — here x at the most
            # referenced columns in shared memory (hub variant); =0 measures the plain walk
            a = sb.matrix_opt(a_plain)
            name += ", matrix_opt (hub columns of x in shared memory)"
        x = G.dense_uniform((n,), 5, torch.float32, dev)
        y = torch.empty(m, device=dev)
        t0 = time.perf_counter()
        info = sb.multiply_inspect(a, x, y)
        torch.cuda.synchronize()
        extra["inspect_ms"] = (time.perf_counter() - t0) * 1e3
        extra["max_row_len"] = info.max_row_len
        extra["empty_rows"] = info.empty_rows
        t0 = time.perf_counter()
        sb.multiply_execute(info, a, x, y)           # first product: builds the lazy tables
        torch.cuda.synchronize()
        extra["first_execute_ms"] = (time.perf_counter() - t0) * 1e3
        extra["spmv_variant"] = info.spmv_variant
        extra["hub_columns"] = info.hub_count
        extra["hub_reference_share"] = info.hub_refs / max(nnz, 1)

        def fn(i):
            sb.multiply_execute(info, a, x, y)

        def plain_walk_ms():
            """the same product without matrix_opt (warp-stream kernel), for the A/B"""
            y2 = torch.empty_like(y)
            info2 = sb.multiply_inspect(a_plain, x, y2)
            ms2 = _time_loop(lambda i: sb.multiply_execute(info2, a_plain, x, y2), K, W)
            same = bool(torch.equal(y, y2))
            v2 = info2.spmv_variant
            info2.close()
            return {"ms": ms2, "spmv_variant": v2, "bit_identical_to_headline": same}
        side_measurements = {"plain_walk": plain_walk_ms} if a is not a_plain else {}
        flops, nbytes, dtype = 2.0 * nnz, _bytes_spmv(nnz, m, n, 4), "f32"
        extra["l2_policy"] = "inputs larger than L2 (2.3 GB)"
        launches_of = lambda: info.total_launches
        cmp_args = ("spmv", [(m, n, rp, ci, v, x, torch.empty_like(y))], 1.0)
        result_of = lambda: y
        probe_ops = [(ci, v, x)]
        cpu_args = ("spmv", m, n, rp, ci, v, x, None, m, 2.0, "C4 full product")
    else:
        k = 32 if wl == "c3k32" else 128
        m = n = 2_000_000
        v, rp, ci, shape = G.uniform_random_csr(m, n, 16, seed=3, dtype=torch.float32, device=dev)
        nnz = int(ci.numel())
        a = sb.csr_view(v, rp, ci, shape, nnz)
        B = G.dense_uniform((n, k), 4, torch.float32, dev)
        C = torch.empty((m, k), device=dev)
        info = sb.multiply_inspect(a, B, C)

        def fn(i):
            sb.multiply_execute(info, a, B, C)
        flops, nbytes, dtype = 2.0 * nnz * k, _bytes_spmm(nnz, m, n, k, 4), "f32"
        name = f"C3 CSR SpMM fp32 2M x 2M, 16 nnz/row, row-major B k={k}"
        extra["l2_policy"] = "inputs larger than L2"
        extra["gather_model_bytes"] = nnz * 8 + (m + 1) * 4 + nnz * k * 4 + m * k * 4
        launches_of = lambda: info.total_launches
        cmp_args = ("spmm", [(m, n, rp, ci, v, B, torch.empty_like(C))], 1.0)
        result_of = lambda: C
        cpu_args = ("spmm", m, n, rp, ci, v, B, None, 400_000 if k == 32 else 100_000, 2.0 * k,
                    f"C3 k={k} row block")

    sampler.start()
    l0 = launches_of()
    ms = _time_loop(fn, K, W)
    launches = launches_of() - l0
    t_load = time.perf_counter()            # keep the same load up for the 100 ms sampler
    while time.perf_counter() - t_load < 0.6:
        for i in range(20):
            fn(i)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "the timed region plus 0.6 s of the same loop right after it"
    for key, measure in locals().get("side_measurements", {}).items():
        try:
            extra[key] = measure()
        except Exception as exc:                       # a side number never costs the line
            extra[key] = {"error": repr(exc)}
    achieved = nbytes / (ms * 1e-3) / 1e9
    from bench import ncu_traffic
    traffic = ncu_traffic("c4_plain_walk" if wl == "c4" and extra.get("spmv_variant") != 3 else wl)
    if wl.startswith("c3"):
        extra["spmm_variant"] = info.spmm_variant
    # same-box vendor comparator (the reference's NVIDIA backend is a cuSPARSE wrapper)
    cusparse = cusparse_compare(cmp_args[0], cmp_args[1], K, cmp_args[2],
                                transpose=locals().get("cmp_transpose", False))
    if cusparse and "unavailable" not in cusparse:
        ours, theirs = result_of(), cmp_args[1][0][6]
        fn(0)                                          # operand set 0 again
        torch.cuda.synchronize()
        err = (ours.double() - theirs.double()).abs().max().item()
        ref = theirs.double().abs().max().item()
        cusparse["max_abs_diff_vs_ours"] = err
        cusparse["max_abs_result"] = ref
        best = min(vv for kk, vv in cusparse.items() if kk.startswith("CUSPARSE_"))
        cusparse["ours_over_best_cusparse"] = best / ms
    gather = None
    if not wl.startswith("c3") and probe_ops is not None:
        from spblas_reference_b200 import _cabi
        pms = gather_ceiling(_cabi, probe_ops, K)
        gather = {"probe_ms": pms, "frac_of_probe": pms / ms,
                  "gathers_per_ns": nnz / (pms * 1e6),
                  "what": "csrc/probe.cu on this workload's colind/values/x: 128-bit streaming "
                          "loads + one gather of x per nonzero, no rows, no reduction — the "
                          "ceiling any SpMV kernel has on this index stream (the L1 tag stage "
                          "takes one 128-byte line per cycle per SM: 148 x 1.965 GHz = 291 "
                          "gathers/ns)"}
    line = {
        "metric": "CSR SpMM GFLOP/s" if wl.startswith("c3") else "CSR SpMV GFLOP/s",
        "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": 1, "steps": K,
        "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": dict(workload=name, nnz=nnz, **extra),
        "gbs": achieved,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                     # what the launch really moved through DRAM (ncu), per second: for the
                     # gather-bound configs THIS is what sits against the HBM peak
                     "dram_gbs_at_ncu_traffic": (traffic / (ms * 1e-3) / 1e9) if traffic else None,
                     "gather_ceiling": gather},
        "clocks": clocks, "gpu_launches": int(launches),
        "cusparse": cusparse,
        "cpu_baseline": None if (args.no_cpu_baseline or cpu_args is None) else cpu_baseline(*cpu_args),
    }
    print(json.dumps(line), flush=True)


def run_transpose(args, sb, G, dev, peak, peak_src, sampler):
    """--workload t1 | t4: transpose_inspect + transpose (CSR -> CSR) of the C1
    matrix (uniform random, 1M x 1M, 10 per row) or of an R-MAT (scale 22, edge factor 16),
    fp32 / int32.  A step is the execute phase transpose(info, a, b) (structure copied out,
    values moved through the permutation); the inspect phase (the sort) is timed beside it."""
    import numpy as np
    K, W = max(1, args.steps), max(3, args.warmup)
    if args.workload == "t1":
        v, rp, ci, shape = G.uniform_random_csr(1_000_000, 1_000_000, 10, seed=0,
                                                dtype=torch.float32, device=dev)
        name = "transpose of the C1 matrix (uniform random 1M x 1M, 10 nnz/row) fp32/int32"
    else:
        v, rp, ci, shape = G.rmat_csr(22, 16, seed=24, dtype=torch.float32, device=dev)
        name = "transpose of an R-MAT scale 22 (edge factor 16) fp32/int32"
    m, n = shape
    nnz = int(ci.numel())
    a = sb.csr_view(v, rp, ci, shape, nnz)
    b = sb.csr_view(torch.empty(nnz, device=dev), torch.empty(n + 1, dtype=torch.int32, device=dev),
                    torch.empty(nnz, dtype=torch.int32, device=dev), (n, m), 0)
    info = sb.transpose_inspect(a, b)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):                       # re-inspect into the same info: buffers reused
        sb.transpose_inspect(info, a, b)
        torch.cuda.synchronize()
    inspect_ms = (time.perf_counter() - t0) / 3 * 1e3
    sampler.start()
    ms = _time_loop(lambda i: sb.transpose(info, a, b), K, W)
    clocks = sampler.stop()
    # compulsory bytes of B = A^T: read A once, write B once
    nbytes = nnz * 8 + (m + 1) * 4 + nnz * 8 + (n + 1) * 4
    # what the execute phase moves: structure of B copied (read + write), permutation and
    # values read, values written
    moved = 2 * (nnz * 4 + (n + 1) * 4) + nnz * 4 + nnz * 4 + nnz * 4
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        vh, rph, cih = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy()
        impl = "reference" if O.have_ref() else "oracle"
        t0 = time.perf_counter()
        want = O.transpose(shape, rph, cih, vh, impl=impl)
        sec = time.perf_counter() - t0
        same = (np.array_equal(want[0], b.values.cpu().numpy()) and
                np.array_equal(want[1], b.rowptr.cpu().numpy()) and
                np.array_equal(want[2], b.colind.cpu().numpy()))
        cpu = {"value": nbytes / sec / 1e9, "unit": "GB/s", "cores": 1,
               "kind": "reference" if impl == "reference" else "port", "seconds": sec,
               "sample": "the full transpose, once (the reference's transpose is serial)",
               "bit_identical_to_gpu_result": bool(same), "host_cores_available": os.cpu_count()}
    line = {
        "metric": "CSR transpose GB/s", "value": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "nnz": nnz, "inspect_ms": inspect_ms,
                   "inspect_plus_execute_ms": inspect_ms + ms,
                   "l2_policy": "rotating is not needed: 10M+ entries x 20 B exceed L2"},
        "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": nbytes,
                     "bytes_moved_by_execute_phase": moved, "peak_source": peak_src,
                     "note": "the values travel through a permutation: one scattered 4-byte read "
                             "per entry, the same L1->L2 request-port bound as SpMV's gathers"},
        "clocks": clocks, "gpu_launches": 1 * K, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def run_trsv(args, sb, G, dev, peak, peak_src, sampler):
    """--workload trsv: triangular_solve on the lower triangle (explicit diagonal) of the C2
    matrix — 5-point Poisson on a 4096 x 4096 grid, fp64/int32: 16.7 M unknowns in 8191 level
    sets (the anti-diagonals of the grid).  A step is one solve; the level launches are
    replayed from a CUDA graph, with the level-by-level launch time beside it."""
    import numpy as np
    K, W = max(1, args.steps), max(3, args.warmup)
    g = args.grid
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, dev)
    m, nnz = shape[0], int(ci.numel())
    a = sb.csr_view(v, rp, ci, shape, nnz)
    b = G.dense_uniform((m,), 7, torch.float64, dev)
    x = torch.empty(m, dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    info = sb.triangular_solve_inspect(a, sb.lower_triangle, sb.explicit_diagonal, b, x)
    torch.cuda.synchronize()
    inspect_ms = (time.perf_counter() - t0) * 1e3
    solve = lambda i: sb.triangular_solve(info, a, sb.lower_triangle, sb.explicit_diagonal, b, x)
    sampler.start()
    ms = _time_loop(solve, K, W)
    clocks = sampler.stop()
    os.environ["SPBLAS_B200_TRSV_GRAPH"] = "0"
    info0 = sb.triangular_solve_inspect(a, sb.lower_triangle, sb.explicit_diagonal, b, x)
    os.environ.pop("SPBLAS_B200_TRSV_GRAPH")
    x0 = torch.empty_like(x)
    ms_direct = _time_loop(lambda i: sb.triangular_solve(info0, a, sb.lower_triangle,
                                                         sb.explicit_diagonal, b, x0), max(3, K // 4), 2)
    same_paths = bool(torch.equal(x, x0))
    used = (nnz + m) // 2                       # stored entries of the lower triangle incl. diagonal
    flops = 2.0 * used
    nbytes = nnz * 12 + (m + 1) * 4 + 3 * m * 8   # whole rows are read; b read, x read and written
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        impl = "reference" if O.have_ref() else "oracle"
        vh, rph, cih, bh = v.cpu().numpy(), rp.cpu().numpy(), ci.cpu().numpy(), b.cpu().numpy()
        t0 = time.perf_counter()
        want = O.trsv(m, rph, cih, vh, bh, upper=False, unit=False, impl=impl)
        sec = time.perf_counter() - t0
        cpu = {"value": flops / sec / 1e9, "unit": "GFLOP/s", "cores": 1,
               "kind": "reference" if impl == "reference" else "port", "seconds": sec,
               "sample": "the full solve, once (the reference's triangular_solve is serial)",
               "bit_identical_to_gpu_result": bool(np.array_equal(want, x.cpu().numpy())),
               "host_cores_available": os.cpu_count()}
    line = {
        "metric": "CSR SpTRSV GFLOP/s", "value": flops / (ms * 1e-3) / 1e9, "unit": "GFLOP/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"lower-triangular solve (explicit diagonal) with the C2 matrix, "
                               f"poisson2d {g}x{g}, fp64/int32", "rows": m, "nnz": nnz,
                   "levels": info.trsv_levels, "inspect_ms": inspect_ms,
                   "inspect_sweeps": info.trsv_sweeps,
                   "ms_per_step_level_by_level_launches": ms_direct,
                   "graph_and_direct_bit_identical": same_paths,
                   "us_per_level": ms * 1e3 / max(info.trsv_levels, 1)},
        "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                     "note": "latency-bound by construction: 8191 dependent levels of <= 4096 rows; "
                             "the figure of merit is microseconds per level, not bytes per second"},
        "clocks": clocks, "gpu_launches": int(info.trsv_levels) * K, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def run_c5(args, sb, G, dev, peak, peak_src, sampler, world, rank, barrier, max_over_ranks,
           sum_over_ranks):
    """C5: R-MAT (edge factor 16) CSR SpMV fp64 with int32 indices and int64 offsets,
    nnz-balanced row blocks over the ranks, x replicated, iterated y -> x with an allgather
    of the blocks (sharded.py picks it: an R-MAT block references almost every column).
    Weak-scaled by default: scale = 24 + log2(N) (scale 27 at N = 8, as BASELINE.json)."""
    import math
    from spblas_reference_b200.sharded import ShardedSpMV, balanced_nnz_blocks
    K, W = max(1, args.steps), max(3, args.warmup)
    scale = args.scale if args.scale > 0 else 24 + int(round(math.log2(world)))
    n = 1 << scale
    deg = G.rmat_degrees(scale, 16, seed=27, device=dev)
    rowptr_all = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=rowptr_all[1:])
    blocks = balanced_nnz_blocks(rowptr_all, world)
    max_deg = int(deg.max())
    del deg, rowptr_all
    r0, r1 = blocks[rank]
    v, rp, ci, shape = G.rmat_csr(scale, 16, seed=27, dtype=torch.float64, device=dev,
                                  off_dtype=torch.int64, row_begin=r0, row_end=r1)
    m_loc, nnz_loc = shape[0], int(ci.numel())
    a = sb.csr_view(v, rp, ci, shape, nnz_loc)
    alpha = 1.0 / max_deg                          # keeps the iterates in [0, 1]
    a_scaled = sb.scaled(alpha, a)
    x0 = G.dense_uniform((n,), 5, torch.float64, dev)
    y0 = torch.empty(m_loc, dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    info = sb.multiply_inspect(a, x0, y0)
    torch.cuda.synchronize()
    inspect_ms = (time.perf_counter() - t0) * 1e3
    import os
    op = ShardedSpMV(n, blocks, (0, n), lambda x, y: sb.multiply_execute(info, a_scaled, x, y),
                     torch.float64, dev, info=info,
                     fused=None if os.environ.get("SPBLAS_B200_FUSED", "1") != "0" else False,
                     multicast={"1": True, "0": False}.get(os.environ.get("SPBLAS_B200_MULTICAST")))
    op.set_x(x0)
    del x0, y0
    for _ in range(W):
        op.step()
    barrier()
    if rank == 0:
        sampler.start()
    l0 = info.total_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        op.step()
    e1.record()
    barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1) / K)
    launches = int(sum_over_ranks(info.total_launches - l0))
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    k0.record()
    for _ in range(K):
        op.multiply()
    k1.record()
    barrier()
    kern_ms = max_over_ranks(k0.elapsed_time(k1) / K)
    clocks = sampler.stop() if rank == 0 else None
    total_nnz = int(sum_over_ranks(nnz_loc))
    nbytes = nnz_loc * 12 + (m_loc + 1) * 8 + n * 8 + m_loc * 8   # full replicated x
    achieved = nbytes / (kern_ms * 1e-3) / 1e9
    from spblas_reference_b200 import _cabi
    xp = G.dense_uniform((n,), 5, torch.float64, dev)
    pms = gather_ceiling(_cabi, [(ci, v, xp)], max(5, K // 2))
    del xp
    gather = {"probe_ms": pms, "frac_of_probe": pms / kern_ms,
              "gathers_per_ns": nnz_loc / (pms * 1e6),
              "what": "csrc/probe.cu on rank 0's colind/values and a full x: the same loads as "
                      "SpMV and nothing else — the ceiling of this index stream"}
    if rank == 0:
        line = {
            "metric": "CSR SpMV GFLOP/s", "value": 2.0 * total_nnz / (step_ms * 1e-3) / 1e9,
            "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"C5 R-MAT scale {scale} (edge factor 16) CSR SpMV fp64, int32 "
                                   "indices, int64 offsets, nnz-balanced row blocks, iterated "
                                   "y->x with allgather", "nnz": total_nnz,
                       "rows_rank0": m_loc, "nnz_rank0": nnz_loc, "exchange": op.plan.mode,
                       "exchange_impl": ("fused peer stores" + (" (NVLS multicast)" if getattr(op, "multicast", False) else "")
                                         if op.fused else ("nccl" if world > 1 else "none")),
                       "fused_error": op.fused_error,
                       "kernel_only_ms": kern_ms, "inspect_ms": inspect_ms,
                       "max_row_len": info.max_row_len, "spmv_variant": info.spmv_variant,
                       "l2_policy": "inputs larger than L2"},
            "gbs": achieved,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                         "gather_ceiling": gather,
                         "note": "random 8-byte gathers of x (1 GB at scale 27) cost a 32-byte "
                                 "sector each: the compulsory-bytes roofline is not reachable "
                                 "(SURVEY 8d caveat)"},
            "clocks": clocks, "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)


def run_c5mm(args, sb, G, dev, peak, peak_src, sampler, world, rank, barrier, max_over_ranks,
             sum_over_ranks):
    """C5's SpMM half (BASELINE.json configs[4]): R-MAT (edge factor 16) CSR times a dense
    row-major B with k = 32 in fp64, int32 indices, int64 offsets, nnz-balanced row blocks
    over the ranks, B replicated.  A single product C = A B needs no exchange (SURVEY 8e:
    every rank writes its own block of C), so the step has no collective; time = max over
    ranks.  Weak-scaled: scale = 22 + log2(N) by default (B is k times an x: 8.6 GB at scale
    25); --scale 27 is BASELINE's size (34 GB of B per GPU)."""
    import math
    from spblas_reference_b200.sharded import balanced_nnz_blocks
    K, W = max(1, args.steps), max(3, args.warmup)
    k = 32
    scale = args.scale if args.scale > 0 else 22 + int(round(math.log2(world)))
    n = 1 << scale
    deg = G.rmat_degrees(scale, 16, seed=27, device=dev)
    rowptr_all = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=rowptr_all[1:])
    blocks = balanced_nnz_blocks(rowptr_all, world)
    del deg, rowptr_all
    r0, r1 = blocks[rank]
    v, rp, ci, shape = G.rmat_csr(scale, 16, seed=27, dtype=torch.float64, device=dev,
                                  off_dtype=torch.int64, row_begin=r0, row_end=r1)
    m_loc, nnz_loc = shape[0], int(ci.numel())
    a = sb.csr_view(v, rp, ci, shape, nnz_loc)
    B = G.dense_uniform_rows(n, k, 6, torch.float64, dev)
    C = torch.empty((m_loc, k), dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    info = sb.multiply_inspect(a, B, C)
    torch.cuda.synchronize()
    inspect_ms = (time.perf_counter() - t0) * 1e3

    def fn(i):
        sb.multiply_execute(info, a, B, C)
    for i in range(W):
        fn(i)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = info.total_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        fn(i)
    e1.record()
    barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1) / K)
    launches = int(sum_over_ranks(info.total_launches - l0))
    clocks = None
    if rank == 0:
        t_load = time.perf_counter()        # the same load a little longer for the 100 ms sampler
        while time.perf_counter() - t_load < 0.6:
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks["window"] = "the timed region plus 0.6 s of the same loop on rank 0 right after it"
    barrier()
    total_nnz = int(sum_over_ranks(nnz_loc))
    nbytes = nnz_loc * 12 + (m_loc + 1) * 8 + n * k * 8 + m_loc * k * 8   # B fully replicated
    achieved = nbytes / (step_ms * 1e-3) / 1e9
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline("spmm", m_loc, n, rp, ci, v, B, None, 100_000, 2.0 * k,
                           f"C5 SpMM k={k} row block")
    if rank == 0:
        line = {
            "metric": "CSR SpMM GFLOP/s", "value": 2.0 * total_nnz * k / (step_ms * 1e-3) / 1e9,
            "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"C5 R-MAT scale {scale} (edge factor 16) CSR SpMM fp64, k={k}, "
                                   "int32 indices, int64 offsets, nnz-balanced row blocks, B "
                                   "replicated, single product (no exchange)",
                       "nnz": total_nnz, "rows_rank0": m_loc, "nnz_rank0": nnz_loc,
                       "parallelism": f"rowblock{world}", "exchange": "none",
                       "inspect_ms": inspect_ms, "spmm_variant": info.spmm_variant,
                       "num_segments": info.num_segments, "max_row_len": info.max_row_len,
                       "l2_policy": "inputs larger than L2"},
            "gbs": achieved,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                         "gather_model_bytes": nnz_loc * 12 + (m_loc + 1) * 8 + nnz_loc * k * 8
                         + m_loc * k * 8,
                         "note": "random 256-byte rows of B: the compulsory-bytes roofline is "
                                 "not reachable (SURVEY 8d caveat); the gather model counts one "
                                 "row of B per stored entry"},
            "clocks": clocks, "gpu_launches": launches, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
