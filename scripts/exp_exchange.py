"""Where does the fused exchange's per-step cost go?  (torchrun, N ranks; C2 weak-scaled.)
Times the y -> x loop with: the full fused exchange; the peer stores without the barrier; the
barrier without the peer stores; plain products (no exchange); the NCCL halo exchange.  The
middle two give WRONG iterates (no synchronisation / no data) — they are timed, never checked."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from spblas_reference_b200.sharded import ShardedSpMV, equal_row_blocks

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = 4096
n = g * world * g
blocks = equal_row_blocks(n, world)
r0, r1 = blocks[rank]
v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, dev, r0, r1, gi=g * world)
a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
a_s = sb.scaled(0.125, a)
x0 = G.dense_uniform((n,), 1, torch.float64, dev)
info = sb.multiply_inspect(a, x0, torch.empty(shape[0], dtype=torch.float64, device=dev))
op = ShardedSpMV(n, blocks, (int(ci.min()), int(ci.max()) + 1),
                 lambda x, y: sb.multiply_execute(info, a_s, x, y), torch.float64, dev, info=info, fused=True)
op.set_x(x0)


def timed(step, K=200, W=20):
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(t.item())


out = {"n_gpus": world}
out["fused"] = timed(op.step)


def scatter_only():
    k, ptrs, lo, hi, mc = op._bound[1 - op.cur]
    op._lib.spblas_b200_plan_set_scatter(info._plan, k, ptrs, lo, hi, mc)
    op.local_multiply(op.x[op.cur], op.x[1 - op.cur][op.r0:op.r1])
    op.cur = 1 - op.cur


info.set_barrier((), ())
op._exchange_bound = "manual"
out["peer_stores_no_barrier"] = timed(scatter_only)
info.set_scatter(())
kb, rs, ls = op._bound_barrier
op._lib.spblas_b200_plan_set_barrier(info._plan, kb, rs, ls)


def barrier_only():
    op.local_multiply(op.x[op.cur], op.x[1 - op.cur][op.r0:op.r1])
    op.cur = 1 - op.cur


out["barrier_no_peer_stores"] = timed(barrier_only)
info.set_barrier((), ())
out["plain_products"] = timed(barrier_only)
op._exchange_bound = None
op.fused = False
out["nccl"] = timed(op.step)
if rank == 0:
    print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
dist.destroy_process_group()
