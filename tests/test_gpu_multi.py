"""Two GPUs, one process each (NCCL for the rendezvous only): the FUSED y -> x iteration —
SpMV kernels storing rows into the peer's x replica over NVLink, flag barrier in the carry
fix-up kernel — against the single-process oracle iterate, for the halo plan (Poisson), the
allgather plan (R-MAT, nnz-balanced blocks) and, where the box offers NVLS, the multicast
variant.  Needs >= 2 devices (gpurun --gpus 2); skipped on a one-GPU box."""
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
                     op.fused_error)
    finally:
        dist.destroy_process_group()


def _run(kind, fused, multicast=False):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, fused, multicast, out), nprocs=world,
             join=True)
    assert len(out) == world
    return [out[r] for r in range(world)]


@pytest.mark.parametrize("kind,mode", [("poisson", "halo"), ("rmat", "allgather")])
def test_fused_iteration_two_gpus(kind, mode):
    for got_mode, ok, fused, timeout, epoch, _, err in _run(kind, True):
        assert got_mode == mode and ok, (got_mode, ok, err)
        assert fused and timeout == 0 and epoch == 6


@pytest.mark.parametrize("kind", ["poisson", "rmat"])
def test_nccl_fallback_two_gpus(kind):
    for _, ok, fused, _, epoch, _, _ in _run(kind, False):
        assert ok and not fused and epoch == 0


def test_fused_multicast_two_gpus():
    res = _run("rmat", True, multicast=True)
    for _, ok, fused, timeout, _, mc, err in res:
        assert ok and fused and timeout == 0, err
    if not all(r[5] for r in res):
        pytest.skip("no NVLS multicast on this box: peer stores were used")
