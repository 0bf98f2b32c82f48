"""2, 4 and 8 GPUs, one process each (NCCL for the rendezvous only): the FUSED y -> x iteration —
SpMV kernels storing rows into the peers' x replicas over NVLink, flag barrier in the carry
fix-up kernel — against the single-process oracle iterate, for the halo plan (Poisson: a middle
rank has TWO neighbours from world size 3 on, the case that once scattered every tile), the
allgather plan (R-MAT, nnz-balanced blocks), where the box offers NVLS the multicast variant,
the NCCL fallback, and the automatic choice between the two (fused=None: both are timed at
set-up).  A case needs as many devices as its world size (gpurun --gpus N) and is skipped on a
smaller box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, kind, fused, multicast, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import spblas_reference_b200 as sb
    from spblas_reference_b200 import generators as G
    from spblas_reference_b200.sharded import ShardedSpMV, balanced_nnz_blocks, equal_row_blocks
    from oracle import oracle as O
    from helpers import assert_rows_within_bound
    try:
        if kind == "poisson":
            g = 192
            v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cpu")
            blocks = equal_row_blocks(shape[0], world)
        else:
            v, rp, ci, shape = G.rmat_csr(13, 8, seed=5, dtype=torch.float64, device="cpu")
            blocks = balanced_nnz_blocks(rp, world)
        n = shape[1]
        r0, r1 = blocks[rank]
        rp64 = rp.to(torch.int64)
        k0, k1 = int(rp64[r0]), int(rp64[r1])
        # the shard keeps the global offsets (base != 0) and the global arrays
        a = sb.csr_view(v.to(dev), rp[r0:r1 + 1].to(dev), ci.to(dev), (r1 - r0, n), k1 - k0)
        lci = ci[k0:k1]
        cols = (int(lci.min()), int(lci.max()) + 1) if lci.numel() else (0, 0)
        x0 = G.dense_uniform((n,), 3, torch.float64, dev)
        info = sb.multiply_inspect(a, x0, torch.empty(r1 - r0, dtype=torch.float64, device=dev))
        a_s = sb.scaled(0.125, a)
        op = ShardedSpMV(n, blocks, cols, lambda x, y: sb.multiply_execute(info, a_s, x, y),
                         torch.float64, dev, info=info, fused=fused, multicast=multicast)
        op.set_x(x0)
        ref = x0.cpu().numpy().copy()
        vv, rr, cc = v.numpy(), rp.numpy(), ci.numpy()
        ok, worst = True, 0.0
        for it in range(6):
            op.step()
            torch.cuda.synchronize()
            prev = ref
            ref = O.spmv("csr", shape, rr, cc, vv, prev, alpha_a=0.125)
            bound = 8 * 2.0 ** -52 * O.abs_rowsum(rr, cc, vv, prev, 0.125) * (np.diff(rr) + 2) * (it + 1)
            got = op.x_current.cpu().numpy()
            lo, hi = cols if op.plan.mode == "halo" else (0, n)
            err = np.abs(got[lo:hi] - ref[lo:hi])
            ok &= bool((err <= bound[lo:hi] + 1e-300).all())
            ok &= bool(np.array_equal(op.y_block.cpu().numpy(), got[r0:r1]))
            ref = ref.copy()
            ref[lo:hi] = got[lo:hi]                 # follow the GPU iterate: bound per step
        # a plain product afterwards must not scatter or wait
        op.multiply(exchange=False)
        torch.cuda.synchronize()
        out[rank] = (op.plan.mode, bool(ok), bool(op.fused), int(info.barrier_timeout),
                     int(info.barrier_epoch), bool(getattr(op, "multicast", False)),
                     op.fused_error, op.calibration)
    finally:
        dist.destroy_process_group()


def _run(kind, fused, multicast=False, world=2):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, fused, multicast, out), nprocs=world,
             join=True)
    assert len(out) == world
    return [out[r] for r in range(world)]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind,mode", [("poisson", "halo"), ("rmat", "allgather")])
def test_fused_iteration(kind, mode, world):
    for got_mode, ok, fused, timeout, epoch, _, err, _ in _run(kind, True, world=world):
        assert got_mode == mode and ok, (got_mode, ok, err)
        assert fused and timeout == 0 and epoch == 6


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["poisson", "rmat"])
def test_nccl_fallback(kind, world):
    for _, ok, fused, _, epoch, _, _, _ in _run(kind, False, world=world):
        assert ok and not fused and epoch == 0


@pytest.mark.parametrize("world", [2, 8])
def test_fused_multicast(world):
    res = _run("rmat", True, multicast=True, world=world)
    for _, ok, fused, timeout, _, mc, err, _ in res:
        assert ok and fused and timeout == 0, err
    if not all(r[5] for r in res):
        pytest.skip("no NVLS multicast on this box: peer stores were used")


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["poisson", "rmat"])
def test_automatic_exchange_choice(kind, world):
    """fused=None: both exchanges are timed at set-up, every rank keeps the same one, and the
    iteration is right whichever it is."""
    res = _run(kind, None, world=world)
    kept = {r[7]["kept"] for r in res if r[7]}
    assert len(kept) == 1, res
    for _, ok, fused, timeout, _, _, err, cal in res:
        assert ok and timeout == 0, err
        assert cal is not None and fused == (cal["kept"] == "fused")


def _timeout_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SPBLAS_B200_BARRIER_TIMEOUT_MS"] = "300"          # read when the plan is created
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import spblas_reference_b200 as sb
    from spblas_reference_b200 import generators as G
    from spblas_reference_b200.sharded import ShardedSpMV, equal_row_blocks
    try:
        g = 64
        v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cpu")
        blocks = equal_row_blocks(shape[0], world)
        r0, r1 = blocks[rank]
        k0, k1 = int(rp[r0]), int(rp[r1])
        a = sb.csr_view(v.to(dev), rp[r0:r1 + 1].to(dev), ci.to(dev), (r1 - r0, shape[1]), k1 - k0)
        lci = ci[k0:k1]
        x0 = G.dense_uniform((shape[1],), 3, torch.float64, dev)
        info = sb.multiply_inspect(a, x0, torch.empty(r1 - r0, dtype=torch.float64, device=dev))
        op = ShardedSpMV(shape[1], blocks, (int(lci.min()), int(lci.max()) + 1),
                         lambda x, y: sb.multiply_execute(info, a, x, y), torch.float64, dev,
                         info=info, fused=True)
        op.set_x(x0)
        op.step()                       # both ranks: a good step
        torch.cuda.synchronize()
        dist.barrier()
        raised, msg = False, ""
        if rank == 0:                   # rank 1 never issues its second step
            op.step()                   # rank 0's barrier waits 300 ms, then gives up
            torch.cuda.synchronize()
            try:
                op.step()               # ... and the NEXT execute fails loudly
            except RuntimeError as exc:
                raised, msg = True, str(exc)
        dist.barrier()
        out[rank] = (raised, msg, int(info.barrier_timeout))
    finally:
        dist.destroy_process_group()


def test_fused_barrier_timeout_is_loud():
    """A peer that never arrives: the waiting rank's kernel gives up after
    SPBLAS_B200_BARRIER_TIMEOUT_MS instead of hanging the GPU, and the next execute on the
    plan raises (the step's x was incomplete) — never a silent wrong result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_timeout_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    raised, msg, flag = out[0]
    assert raised and "did not reach the barrier" in msg and flag == 1, out[0]
    assert out[1] == (False, "", 0)
