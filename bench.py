#!/usr/bin/env python
"""bench.py — the measurement contract for the spblas B200 backend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c2|c1|c4|c3k32|c3k128|c5|c5mm|c1t|c4t|t1|t4|trsv]

Workload at N=1 (the configuration BASELINE.json's metric is quoted on, configs[1]):
  C2 — 2D Poisson 5-point stencil on a 4096 x 4096 grid, CSR SpMV in fp64 with int32
  indices/offsets, iterated y -> x with scaled(1/8, A).  A "step" is one
  product x <- (1/8) A x over the whole matrix (1.34 GB of operands, larger than L2, so
  consecutive steps cannot be served from cache; no flush needed).
At N>1 the run is WEAK-scaled: every rank owns one 4096 x 4096 grid's worth of rows of the
(4096 N) x 4096 grid, x is replicated, and the y -> x step exchanges only the halo the
inspect phase found necessary (spblas_reference_b200/sharded.py).  value = GFLOP/s over all
ranks, time = max over ranks (CUDA events).

The same JSON line carries, under "configs", the other BASELINE.json configs measured in the
same run (bench_configs.py): C1, C3 (k = 32, 128) and C4 at N=1, and C5 — R-MAT scale 27, SpMV
and SpMM — STRONG-scaled over the N ranks at every N; each block has its own roofline,
cpu_baseline, e2e and parity.  `--configs none` prints the headline alone, `--configs c1,c4`
a selection.  Every result that is timed is also checked: "parity" compares the device result
with the reference's CPU multiply under the north-star bound.

`--impl reference` times the reference's own CPU multiply on the host (oracle/_ref when it
was built from /root/reference, else the oracle port) on the same workload.

Everything measured goes through the public host API -> C ABI -> sm_100a kernels.  The
oracle is used only for cpu_baseline / --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
ALIGN_STEPS = 8        # untimed steps queued between the host barrier and the start event
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2",
                    choices=["c1", "c2", "c3k32", "c3k128", "c4", "c5", "c5mm", "t1", "t4", "c1t",
                             "c4t", "trsv"])
    ap.add_argument("--scale", type=int, default=0,
                    help="C5 R-MAT scale (default 24 + log2(N): 16.7M rows per GPU; c5mm: "
                         "22 + log2(N))")
    ap.add_argument("--grid", type=int, default=4096, help="C2 grid edge (per GPU)")
    ap.add_argument("--configs", default="all",
                    help="blocks beside the headline: all | none | comma list of c1,c3k32,c3k128,c4,c5 "
                         "(c5 includes its SpMM half; at N>1 only c5 applies)")
    ap.add_argument("--c5-scale", type=int, default=27, help="R-MAT scale of the c5 block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# algorithmic (compulsory) work, SURVEY §8d
# ----------------------------------------------------------------------------------------
def spmv_bytes(nnz, m, n_touched, sT, sI, sO):
    return nnz * (sT + sI) + (m + 1) * sO + n_touched * sT + m * sT


def spmm_bytes(nnz, m, n, k, sT, sI, sO):
    return nnz * (sT + sI) + (m + 1) * sO + n * k * sT + m * k * sT


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(workload)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = None                       # created by start(): only the rank that samples has a file
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
        try:
            with os.fdopen(fd, "w") as out:    # the child keeps its own copy of the descriptor
                self.proc = subprocess.Popen(
                    ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                     "--format=csv,noheader,nounits", "-lms", "100"],
                    stdout=out, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout_s=4.0):
        """Block until nvidia-smi has written its first line: its start-up (NVML attaches to
        every GPU of the box) must not fall into the timed region — on an 8-GPU box it cost the
        20 timed steps 25 us each."""
        if self.proc is None:
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < timeout_s:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            if self.path:
                try:
                    os.unlink(self.path)
                except OSError:
                    pass
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smmax.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smmax),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------
# reference arm: the reference's own CPU multiply on the host
# ----------------------------------------------------------------------------------------
def host_matrix_c2(g, gi, r0, r1):
    import torch
    from spblas_reference_b200 import generators as G
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cpu", r0, r1, gi=gi)
    return v.numpy(), rp.numpy(), ci.numpy(), shape


def cpu_reference_spmv_seconds(v, rp, ci, shape, x, reps, alpha):
    """Times y = alpha A x with the reference's CPU multiply (1 thread: the reference is
    serial).  Returns (best seconds, kind)."""
    from oracle import oracle as O
    O.build()
    impl, kind = ("reference", "reference") if O.have_ref() else ("oracle", "port")
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        O.spmv("csr", shape, rp, ci, v, x, alpha_a=alpha, impl=impl)
        best = min(best, time.perf_counter() - t0)
    return best, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    g = args.grid
    v, rp, ci, shape = host_matrix_c2(g, g, 0, g * g)
    m, n = shape
    nnz = len(ci)
    import torch
    from spblas_reference_b200 import generators as G
    x = G.dense_uniform((n,), 1, torch.float64, "cpu").numpy()      # the b200 arm's x0
    from oracle import oracle as O
    O.build()
    impl, kind = ("reference", "reference") if O.have_ref() else ("oracle", "port")
    for _ in range(max(1, min(args.warmup, 3))):
        y = O.spmv("csr", shape, rp, ci, v, x, alpha_a=0.125, impl=impl)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        y = O.spmv("csr", shape, rp, ci, v, x, alpha_a=0.125, impl=impl)
        x = y
    dt = (time.perf_counter() - t0) / steps
    gflops = 2.0 * nnz / dt / 1e9
    line = {
        "impl": "reference", "metric": "CSR SpMV GFLOP/s", "value": gflops, "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"C2 poisson2d {g}x{g} CSR SpMV fp64/int32, iterated y->x, "
                               "alpha=1/8 via scaled(); one full product per step",
                   "rows": m, "nnz": nnz,
                   "sample_of": (f"the {args.gpus}-GPU arm's weak-scaled workload ({g * args.gpus}x{g} "
                                 f"grid): one GPU's {g}x{g} block of rows per step — a rate "
                                 "(GFLOP/s) of a serial code does not depend on how many blocks "
                                 "it is given") if args.gpus > 1 else "the 1-GPU arm's full workload"},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                         "sample": "the full product, every step (reference CPU multiply is "
                                   "serial: 1 thread)",
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gbs": spmv_bytes(nnz, m, n, 8, 4, 4) / dt / 1e9,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import spblas_reference_b200 as sb
    from spblas_reference_b200 import generators as G
    from spblas_reference_b200.sharded import ShardedSpMV, equal_row_blocks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    K, W = max(1, args.steps), max(3, args.warmup)
    peak, peak_src = measured_peak()

    if args.workload == "c5":
        from bench_extra import run_c5   # R-MAT fp64 / int64 offsets, nnz-balanced row blocks
        run_c5(args, sb, G, dev, peak, peak_src, ClockSampler(local_rank), world, rank,
               barrier, max_over_ranks, sum_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "c5mm":
        from bench_extra import run_c5mm  # R-MAT fp64 SpMM, row blocks, B replicated, no exchange
        run_c5mm(args, sb, G, dev, peak, peak_src, ClockSampler(local_rank), world, rank,
                 barrier, max_over_ranks, sum_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "trsv":
        from bench_extra import run_trsv        # triangular solve
        run_trsv(args, sb, G, dev, peak, peak_src, ClockSampler(local_rank))
        return
    if args.workload in ("t1", "t4"):
        from bench_extra import run_transpose   # CSR -> CSR transpose
        run_transpose(args, sb, G, dev, peak, peak_src, ClockSampler(local_rank))
        return
    if args.workload != "c2":
        from bench_extra import run_extra   # single-GPU side workloads (C1, C3, C4)
        run_extra(args, sb, G, dev, peak, peak_src, ClockSampler(local_rank))
        return

    # ---- C2: one g x g grid of rows per rank --------------------------------------------
    g = args.grid
    gi = g * world
    n = gi * g
    blocks = equal_row_blocks(n, world)
    r0, r1 = blocks[rank]
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, dev, r0, r1, gi=gi)
    m_loc, nnz_loc = shape[0], int(ci.numel())
    a = sb.csr_view(v, rp, ci, shape, nnz_loc)
    a_scaled = sb.scaled(0.125, a)
    cmin, cmax = int(ci.min()), int(ci.max())

    # x0 = U[0,1) (splitmix64, seed 1: SURVEY 8d's variant).  With x0 = 1 the iterate is 0 on
    # every interior row from the second step on, and the timed loop would multiply zeros.
    x0 = G.dense_uniform((n,), 1, torch.float64, dev)
    if os.environ.get("SPBLAS_B200_BENCH_X0") == "ones":     # round 1's operand, for A/B runs only
        x0.fill_(1.0)
    info = sb.multiply_inspect(a, x0, torch.empty(m_loc, dtype=torch.float64, device=dev))
    t_ins0 = time.perf_counter()
    sb.multiply_inspect(info, a, x0, torch.empty(m_loc, dtype=torch.float64, device=dev))
    torch.cuda.synchronize()
    inspect_ms = (time.perf_counter() - t_ins0) * 1e3

    op = ShardedSpMV(n, blocks, (cmin, cmax + 1),
                     lambda x, y: sb.multiply_execute(info, a_scaled, x, y),
                     torch.float64, dev, info=info,
                     fused=None if os.environ.get("SPBLAS_B200_FUSED", "1") != "0" else False)
    op.set_x(x0)

    # ---- warm-up, then the timed region ---------------------------------------------------
    for _ in range(W):
        op.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # N > 1: the ranks leave the host barrier tens of microseconds apart, and with the exchange
    # fused into the kernels the first step of every rank waits for the last rank's launch — a
    # one-off that 20 steps of 0.23 ms would carry as "per step".  ALIGN untimed steps are
    # queued behind the barrier and ahead of the start event (no host synchronisation in
    # between): the in-kernel flag barrier brings the GPUs into lock-step before the clock starts.
    # N = 1 runs them too: the GPU has just idled while the clock sampler started (up to a
    # second), and the first steps after an idle period run below the clocks of the loop.
    align = ALIGN_STEPS
    for _ in range(align):
        op.step()
    launches0 = info.total_launches
    e0.record()
    for _ in range(K):
        op.step()
    e1.record()
    barrier()
    step_ms = max_over_ranks(e0.elapsed_time(e1) / K)
    launches = int(sum_over_ranks(info.total_launches - launches0))

    # kernel-only loop (no exchange) for the roofline of the dominant kernel
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    k0.record()
    for _ in range(K):
        op.multiply()
    k1.record()
    barrier()
    kern_ms = max_over_ranks(k0.elapsed_time(k1) / K)
    # nvidia-smi samples every 100 ms and the timed region is ~10 ms: keep the SAME loop
    # running (untimed) until the sampler has seen at least half a second of this load
    # (a step count derived from the all-reduced step time: every rank runs the same
    # number of steps, as the exchange requires).  The iterate decays like (1/8 A)^k: it is
    # re-seeded every 64 steps so that the loop never multiplies denormals or zeros.
    extra_steps = min(20000, int(600.0 / max(step_ms, 1e-3)))
    for i in range(extra_steps):
        if i % 64 == 0:
            op.set_x(x0)
        op.step()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the timed region plus 0.6 s of the same step loop right after it"

    total_nnz = int(sum_over_ranks(nnz_loc))
    flops_step = 2.0 * total_nnz
    value = flops_step / (step_ms * 1e-3) / 1e9
    x_touched = min(n, cmax - cmin + 1)
    bytes_launch = spmv_bytes(nnz_loc, m_loc, x_touched, 8, 4, 4)
    achieved = bytes_launch / (kern_ms * 1e-3) / 1e9

    # ---- parity of what was timed: three more iterations of the (fused) y -> x loop from x0,
    # then (1) every value this rank received is bit for bit the value its owner computed,
    # (2) this rank's rows of the next product against the reference's CPU multiply fed with
    # the same x (per iteration, not compounded).  The CPU product is also the cpu_baseline.
    from bench_configs import check_spmv
    op.set_x(x0)
    del x0
    for _ in range(3):
        op.step()
    x_in = op.x_current.clone()
    replica_ok = True
    if world > 1:
        gathered = [torch.empty(m_loc, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, x_in[r0:r1].contiguous())
        for peer, b0, e0_ in op.plan.recvs:
            pb = blocks[peer][0]
            replica_ok = replica_ok and bool(torch.equal(gathered[peer][b0 - pb:e0_ - pb], x_in[b0:e0_]))
        del gathered
    y_blk = op.step().clone()
    torch.cuda.synchronize()
    barrier_timeout = int(max_over_ranks(float(info.barrier_timeout))) if op.fused else 0
    cpu, parity = None, None
    if not args.no_cpu_baseline:
        cpu, parity = check_spmv(rp, ci, v, x_in, 0.125, y_blk, (0, m_loc), n,
                                 f"the full {g}x{g} product of rank {rank}", reps=3)
        parity["max_err_over_tol"] = max_over_ranks(parity["max_err_over_tol"])
        parity["rows_checked"] = int(sum_over_ranks(parity["rows_checked"]))
        ok = parity["pass"] and replica_ok and barrier_timeout == 0
        parity["pass"] = bool(max_over_ranks(0.0 if ok else 1.0) == 0.0)
        parity["halo_bit_identical_to_owner"] = bool(max_over_ranks(0.0 if replica_ok else 1.0) == 0.0)
        parity["after_iterations"] = 3
        parity["what"] = ("after 3 iterations of the timed y->x loop: every received halo value == "
                          "its owner's (bit for bit), and every rank's rows of the next product "
                          "against the reference's CPU multiply on the same x")
        if not (rank == 0 and world == 1):
            cpu = None                                   # cpu_baseline: rank 0 at N=1 only

    if op.fused:                        # the plan goes back to plain products
        info.set_scatter(())
        info.set_barrier((), ())
    # ---- end to end through the public API with HOST buffers --------------------------------
    e2e = None
    if not args.no_e2e:
        x_host = x_in.cpu().pin_memory()
        y_host = torch.empty(m_loc, dtype=torch.float64).pin_memory()
        x_dev = torch.empty(n, dtype=torch.float64, device=dev)
        y_dev = torch.empty(m_loc, dtype=torch.float64, device=dev)
        ke = max(3, min(K, 10))

        def e2e_step():
            # ONE call of the host-buffer entry point (spblas_b200_spmv_host): the upload
            # of x, the kernels and the download of y are pipelined chunk by chunk inside
            sb.multiply_execute_host(info, a_scaled, x_host, y_host)
            torch.cuda.current_stream().synchronize()

        def e2e_step_serial():
            x_dev.copy_(x_host, non_blocking=True)          # H2D of the step's input
            sb.multiply_execute(info, a_scaled, x_dev, y_dev)
            y_host.copy_(y_dev, non_blocking=True)          # D2H of the step's result
            torch.cuda.current_stream().synchronize()

        def time_e2e(step):
            for _ in range(2):
                step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                step()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / ke * 1e3)

        serial_ms = time_e2e(e2e_step_serial)
        y_serial = y_host.clone()
        e2e_ms = time_e2e(e2e_step)
        same = bool(torch.equal(y_serial, y_host)) and bool(torch.equal(y_host, y_blk.cpu()))
        e2e = {"value": flops_step / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": int(n * 8) * world, "d2h_bytes_per_step": int(m_loc * 8) * world,
               "ms_per_step": e2e_ms, "steps": ke,
               "what": "multiply_execute_host: pinned host x -> device, SpMV kernels, y -> pinned "
                       "host, pipelined over 16 chunks of tiles inside one C-ABI call "
                       "(spblas_b200_spmv_host); A and the inspected plan stay resident "
                       "(operator reuse, as in the y->x loop); every rank moves its own x and y",
               "serial_ms_per_step": serial_ms,
               "serial_what": "the same three steps issued one after the other (copy, "
                              "multiply_execute, copy)",
               "bit_identical_to_device_path": same}
        if parity is not None and not same:
            parity["pass"] = False
        del x_dev, y_dev, x_host, y_host, y_serial

    # ---- same-box vendor comparator: cuSPARSE as the reference's NVIDIA backend calls it ----
    cusparse = None
    if rank == 0 and world == 1:
        from bench_extra import cusparse_compare
        yc = torch.empty(m_loc, dtype=torch.float64, device=dev)
        cusparse = cusparse_compare("spmv", [(m_loc, n, rp, ci, v, x_in, yc)], K, 0.125)
        if cusparse and "unavailable" not in cusparse:
            yo = torch.empty_like(yc)
            sb.multiply_execute(info, a_scaled, x_in, yo)
            torch.cuda.synchronize()
            cusparse["max_abs_diff_vs_ours"] = (yo - yc).abs().max().item()
            best = min(vv for kk, vv in cusparse.items() if kk.startswith("CUSPARSE_"))
            cusparse["ours_over_best_cusparse"] = best / kern_ms
            del yo
        del yc

    line = None
    if rank == 0:
        line = {
            "metric": "CSR SpMV GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"C2 poisson2d {g}x{g} per GPU ({gi}x{g} global) CSR SpMV fp64/int32, "
                            "iterated y->x, alpha=1/8 via scaled(); one full product per step",
                "rows_per_gpu": m_loc, "nnz_per_gpu": nnz_loc, "parallelism": f"rowblock{world}",
                "x0": "U[0,1), splitmix64 seed 1",
                "exchange": op.plan.mode, "halo_elems_per_step": op.plan.recv_elems,
                "exchange_impl": op.exchange_impl if world > 1 else "none",
                "exchange_calibration": getattr(op, "calibration", None),
                "l2_policy": "inputs larger than L2 (1.34 GB per product), no flush",
                "timed_region": (f"host barrier + synchronize, {align} untimed alignment steps queued "
                                 f"on the stream, start event, {K} steps, stop event, synchronize + barrier"),
                "inspect_ms": inspect_ms,
            },
            "gbs": bytes_launch * world / (step_ms * 1e-3) / 1e9,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic("c2"),
                "kernel": "spmv_pipe_kernel<double,int,int,8> (+ carry fix-up kernel, ~2% of the step)",
                "algorithmic_bytes_per_launch": bytes_launch, "kernel_ms": kern_ms,
                "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0,
            },
            "clocks": clocks,
            "gpu_launches": launches,
            "e2e": e2e,
            "cpu_baseline": cpu,
            "parity": parity,
            "cusparse": cusparse,
        }

    # ---- the other configs, in the same line -------------------------------------------------
    del op, x_in, y_blk, a, a_scaled, v, rp, ci
    info.close()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    wanted = {"all": ["c1", "c3k32", "c3k128", "c4", "c5"], "none": []}.get(
        args.configs, [t for t in args.configs.split(",") if t])
    if world > 1:
        wanted = [t for t in wanted if t == "c5"]        # the single-GPU configs are N=1 blocks
    configs, errors = {}, {}
    if wanted:
        import bench_configs as BC
        ctx = {"sb": sb, "G": G, "dev": dev, "K": K, "W": W, "peak": peak, "peak_src": peak_src,
               "traffic": ncu_traffic, "world": world, "rank": rank, "barrier": barrier,
               "max": max_over_ranks, "sum": sum_over_ranks}
        runners = {"c1": lambda: {"c1": BC.config_c1(ctx)},
                   "c3k32": lambda: {"c3k32": BC.config_c3(ctx, 32)},
                   "c3k128": lambda: {"c3k128": BC.config_c3(ctx, 128)},
                   "c4": lambda: {"c4": BC.config_c4(ctx)},
                   "c5": lambda: BC.config_c5(ctx, args.c5_scale)}
        for name in wanted:
            t0 = time.perf_counter()
            try:
                blocks_out = runners[name]()
                for key, blk in blocks_out.items():
                    blk["wall_s"] = time.perf_counter() - t0
                    configs[key] = blk
            except Exception as exc:               # a side block never costs the headline
                if world > 1:
                    raise                          # (ranks must not diverge inside collectives)
                errors[name] = repr(exc)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    if rank == 0:
        line["configs"] = configs
        if errors:
            line["config_errors"] = errors
        line["parity_all_pass"] = bool((parity or {}).get("pass", False) and
                                       all(b.get("parity", {}).get("pass", False) for b in configs.values()))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
