"""Row-block sharding of one SpMV across the GPUs of a node (one process per GPU,
torch.distributed for the plumbing: NCCL over NVLink on the GPU box, gloo in CPU tests).

The reference has no multi-GPU path (no NCCL/MPI anywhere); this module is the
new code of SURVEY §8e.  Rows are independent — y[rows_p] = A[rows_p, :] x — so a single
product needs NO communication: every rank multiplies its row block against its replica of
x.  Only the iterative y -> x use (notes/spmv.hpp:19-23 in the reference) has an exchange
step: after each product a rank needs the entries of the new x that its columns reference.
The inspect phase records the column interval [cmin, cmax] its block touches and picks
  * "halo":      point-to-point exchange of just the overlap with each peer's row block
                 (banded matrices: 5-point Poisson needs g entries per neighbour), or
  * "allgather": every block to every rank (R-MAT and other unstructured matrices).

On GPUs the exchange can be FUSED into the product (`fused=True`; with `fused=None`, the
default, both ways are timed for a few steps at set-up on the operator's own matrix and the
faster one is kept — every rank takes the same decision): both x replicas live in symmetric memory
(torch.distributed._symmetric_memory supplies the allocation and the peer mapping — plumbing),
the SpMV kernels store every row a peer needs straight into that peer's next-x replica, and
the carry fix-up kernel ends with the cross-GPU flag barrier (include/spblas_b200.h:
spblas_b200_plan_set_scatter / _set_barrier).  An iteration is then the same two kernel
launches as on one GPU and contains no collective call.  The NCCL path below stays as the
fallback (and is what the gloo CPU tests drive).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def equal_row_blocks(m: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of ceil(m / world) rows (last ones may be short or empty)."""
    per = (m + world - 1) // world
    return [(min(m, p * per), min(m, (p + 1) * per)) for p in range(world)]


def balanced_nnz_blocks(rowptr: torch.Tensor, world: int) -> List[Tuple[int, int]]:
    """Contiguous row blocks holding ~nnz/world entries each: boundary p is the first row
    whose offset reaches p * nnz / world (binary search in rowptr)."""
    rp = rowptr.to(torch.int64)
    m = rp.numel() - 1
    base, nnz = int(rp[0]), int(rp[-1] - rp[0])
    targets = torch.tensor([base + (p * nnz) // world for p in range(1, world)],
                           dtype=torch.int64, device=rp.device)
    cuts = torch.searchsorted(rp, targets, right=False).clamp_(0, m).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [m]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[p], bounds[p + 1]) for p in range(world)]


@dataclass
class ExchangePlan:
    mode: str                                   # "none" | "halo" | "allgather"
    sends: List[Tuple[int, int, int]]           # (peer, begin, end) global rows I send (from my block)
    recvs: List[Tuple[int, int, int]]           # (peer, begin, end) global rows I receive
    blocks: List[Tuple[int, int]]               # row block of every rank
    recv_elems: int = 0


def plan_exchange(blocks: Sequence[Tuple[int, int]], needs: Sequence[Tuple[int, int]],
                  rank: int, halo_fraction: float = 0.5) -> ExchangePlan:
    """needs[p] = half-open column interval rank p's block references (empty: (0, 0)).
    Pure function of replicated metadata, so every rank derives matching sends/recvs."""
    world = len(blocks)
    n_remote_total = 0
    pair = {}
    for dst in range(world):
        lo, hi = needs[dst]
        for src in range(world):
            if src == dst:
                continue
            b, e = max(lo, blocks[src][0]), min(hi, blocks[src][1])
            if b < e:
                pair[(src, dst)] = (b, e)
                if dst == rank:
                    n_remote_total += e - b
    if world == 1:
        return ExchangePlan("none", [], [], list(blocks))
    total_rows = blocks[-1][1]
    # one global decision: halo only if EVERY rank needs a small part of the remote rows
    worst = 0.0
    for dst in range(world):
        need = sum(e - b for (s, d), (b, e) in pair.items() if d == dst)
        remote = total_rows - (blocks[dst][1] - blocks[dst][0])
        worst = max(worst, need / remote if remote > 0 else 0.0)
    if worst > halo_fraction:
        recvs = [(p, blocks[p][0], blocks[p][1]) for p in range(world) if p != rank]
        sends = [(p, blocks[rank][0], blocks[rank][1]) for p in range(world) if p != rank]
        return ExchangePlan("allgather", sends, recvs, list(blocks),
                            total_rows - (blocks[rank][1] - blocks[rank][0]))
    sends = [(d, b, e) for (s, d), (b, e) in sorted(pair.items()) if s == rank]
    recvs = [(s, b, e) for (s, d), (b, e) in sorted(pair.items()) if d == rank]
    return ExchangePlan("halo", sends, recvs, list(blocks), n_remote_total)


class ShardedSpMV:
    """One rank's part of y = alpha * A x with A row-block sharded.

    `local_multiply(x_full, y_block)` computes this rank's rows (the backend's
    multiply(info, a_local, x, y) on the GPU; the CPU tests pass the oracle).
    x is kept as TWO full-length replicas that ping-pong: the product reads one and writes
    its rows straight into its slice of the other, so the y -> x step costs no local copy.
    """

    def __init__(self, n: int, blocks: Sequence[Tuple[int, int]], col_range: Tuple[int, int],
                 local_multiply: Callable[[torch.Tensor, torch.Tensor], None],
                 dtype, device, group=None, halo_fraction: float = 0.5,
                 info=None, fused: Optional[bool] = None,
                 multicast: Optional[bool] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        assert len(blocks) == self.world
        self.n = n
        self.blocks = list(blocks)
        self.r0, self.r1 = self.blocks[self.rank]
        self.local_multiply = local_multiply
        needs = self._gather_needs(col_range, device)
        self.plan = plan_exchange(self.blocks, needs, self.rank, halo_fraction)
        self.cur = 0
        self.info = info
        self.fused = False
        self.fused_error = None
        self.calibration = None
        self._steps_since_check = 0
        want = fused if fused is not None else True
        if (want and info is not None and self.world > 1 and self.plan.mode != "none"
                and torch.device(device).type == "cuda"):
            try:
                self._setup_fused(n, dtype, device, multicast)
                self.fused = True
            except Exception as exc:                 # no peer mapping on this box: NCCL path
                if fused:
                    raise
                self.fused_error = repr(exc)
        if not self.fused:
            self.x = [torch.zeros(n, dtype=dtype, device=device) for _ in range(2)]
        elif fused is None:
            self._calibrate(device)

    @property
    def exchange_impl(self) -> str:
        if self.world == 1 or self.plan.mode == "none":
            return "none"
        if not self.fused:
            return "nccl " + ("batched send/recv of the halo" if self.plan.mode == "halo" else
                              "allgather (one broadcast per block when the blocks differ)")
        return ("fused: rows stored into the peers' x replicas by the SpMV kernels (" +
                ("one NVLS multimem.st per row" if getattr(self, "multicast", False) else
                 "peer stores over NVLink") + ") + flag barrier in the carry fix-up kernel")

    def _calibrate(self, device, steps: int = 12):
        """fused=None: time `steps` iterations each way on zeros (the kernels' time does not
        depend on the values) and keep the faster exchange; MAX over ranks, so every rank
        decides alike.  The NCCL path runs on the same symmetric buffers."""
        saved = self.cur
        times = {}
        for mode in (True, False):
            self.fused = mode
            for _ in range(3):
                self.step()
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                self.step()
            e1.record()
            torch.cuda.synchronize(device)
            t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            times["fused" if mode else "nccl"] = float(t.item())
        self.fused = times["fused"] <= times["nccl"]
        self.calibration = {"ms_per_step": times, "kept": "fused" if self.fused else "nccl",
                            "steps": steps}
        if not self.fused:
            self.info.set_scatter(())
            self.info.set_barrier((), ())
        self.cur = saved
        for t in self.x:
            t.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    # -- fused exchange: symmetric x replicas, peer destinations, flag barrier ----------
    def _setup_fused(self, n, dtype, device, multicast):
        import torch.distributed._symmetric_memory as symm
        grp = self.group if self.group is not None else dist.group.WORLD
        itemsize = torch.empty((), dtype=dtype).element_size()
        self.x, self._hdl = [], []
        for _ in range(2):
            t = symm.empty(n, dtype=dtype, device=device)
            t.zero_()
            self.x.append(t)
            self._hdl.append(symm.rendezvous(t, grp))
        flags = symm.empty(self.world, dtype=torch.int64, device=device)
        flags.zero_()
        self._flags, self._flags_hdl = flags, symm.rendezvous(flags, grp)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)               # every rank's zeros are in place
        p = self.plan
        peers = sorted({q for q, _, _ in p.sends} | {q for q, _, _ in p.recvs})
        fptr = [int(a) for a in self._flags_hdl.buffer_ptrs]
        self._remote_slots = [fptr[q] + 8 * self.rank for q in peers]
        self._local_slots = [fptr[self.rank] + 8 * q for q in peers]
        self._dsts = []
        for b in range(2):
            ptrs = [int(a) for a in self._hdl[b].buffer_ptrs]
            mc = int(getattr(self._hdl[b], "multicast_ptr", 0) or 0)
            # multicast=None: use NVLS whenever the box offers it for an allgather — one
            # multimem store per row instead of world-1 peer stores (measured at N=8 on
            # C5: 4.28 ms per step against 6.87 ms with peer stores and 7.92 ms with NCCL)
            if (multicast or multicast is None) and mc and p.mode == "allgather":
                # one multimem store per row reaches every GPU's replica (this one too)
                self._dsts.append(([(mc + self.r0 * itemsize, 0, self.r1 - self.r0)], True))
            else:
                self._dsts.append(([(ptrs[q] + self.r0 * itemsize, b0 - self.r0, e0 - self.r0)
                                    for q, b0, e0 in p.sends], False))
        self.multicast = bool(self._dsts[0][1])
        # the C-ABI arguments of both ping-pong states, built once (a step then costs two
        # plain foreign calls, no Python list handling)
        import ctypes as C
        from . import _cabi
        self._lib = _cabi.lib()
        self._bound = []
        for dsts, mc in self._dsts:
            k = len(dsts)
            self._bound.append((k, (C.c_void_p * max(k, 1))(*[int(d[0]) for d in dsts]),
                                (C.c_int64 * max(k, 1))(*[int(d[1]) for d in dsts]),
                                (C.c_int64 * max(k, 1))(*[int(d[2]) for d in dsts]), 1 if mc else 0))
        k = len(peers)
        self._bound_barrier = (k, (C.c_void_p * max(k, 1))(*self._remote_slots),
                               (C.c_void_p * max(k, 1))(*self._local_slots))
        self._exchange_bound = None              # which ping-pong state the plan holds now

    def _gather_needs(self, col_range, device):
        if self.world == 1:
            return [tuple(col_range)]
        mine = torch.tensor(list(col_range), dtype=torch.int64, device=device)
        out = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return [(int(t[0]), int(t[1])) for t in out]

    # -- replicated operand -----------------------------------------------------------
    def set_x(self, x_full: torch.Tensor):
        self.x[self.cur].copy_(x_full)

    @property
    def x_current(self) -> torch.Tensor:
        return self.x[self.cur]

    @property
    def y_block(self) -> torch.Tensor:
        """The rows this rank produced last (a slice of the current x replica)."""
        return self.x[self.cur][self.r0:self.r1]

    # -- one product, no communication ---------------------------------------------------
    def multiply(self, exchange: bool = False) -> torch.Tensor:
        nxt = self.x[1 - self.cur]
        y = nxt[self.r0:self.r1]
        if self.fused:
            want = (1 - self.cur) if exchange else None
            if want != self._exchange_bound:
                if exchange:
                    k, ptrs, lo, hi, mc = self._bound[want]
                    st = self._lib.spblas_b200_plan_set_scatter(self.info._plan, k, ptrs, lo, hi, mc)
                    if st == 0 and self._exchange_bound is None:
                        kb, rs, ls = self._bound_barrier
                        st = self._lib.spblas_b200_plan_set_barrier(self.info._plan, kb, rs, ls)
                    if st != 0:
                        raise RuntimeError("fused exchange: " + self.info._err())
                else:
                    self.info.set_scatter(())
                    self.info.set_barrier((), ())
                self._exchange_bound = want
        elif getattr(self, "_exchange_bound", None) is not None:
            self.info.set_scatter(())            # calibration left the plan in a fused state
            self.info.set_barrier((), ())
            self._exchange_bound = None
        self.local_multiply(self.x[self.cur], y)
        return y

    # -- y -> x --------------------------------------------------------------------------
    def exchange(self):
        """Make the rows other ranks produced visible in the next x replica, then flip."""
        nxt = self.x[1 - self.cur]
        p = self.plan
        if self.fused:
            self.cur = 1 - self.cur                  # the kernels already did it
            return
        if p.mode == "allgather":
            sizes = {e - b for (b, e) in p.blocks}
            if len(sizes) == 1 and self.n == self.world * (self.r1 - self.r0):
                dist.all_gather_into_tensor(nxt, nxt[self.r0:self.r1].clone(), group=self.group)
            else:
                works = [dist.broadcast(nxt[b:e], src=self._global(src), group=self.group,
                                        async_op=True)
                         for src, (b, e) in enumerate(p.blocks) if e > b]
                for w in works:
                    w.wait()
        elif p.mode == "halo":
            ops = []
            for peer, b, e in p.sends:
                ops.append(dist.P2POp(dist.isend, nxt[b:e], self._global(peer), self.group))
            for peer, b, e in p.recvs:
                ops.append(dist.P2POp(dist.irecv, nxt[b:e], self._global(peer), self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        self.cur = 1 - self.cur

    def _global(self, group_rank: int) -> int:
        if self.group is None:
            return group_rank
        return dist.get_global_rank(self.group, group_rank)

    def step(self) -> torch.Tensor:
        """x <- alpha * A x (one iteration of the y -> x loop)."""
        y = self.multiply(exchange=True)
        self.exchange()
        return y
