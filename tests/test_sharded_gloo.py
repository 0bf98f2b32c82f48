"""Row-block sharding over torch.distributed with the gloo backend, world_size 2, 3 and 4, on CPU:
the partitioners, the halo / allgather exchange plan and the ping-pong y -> x step.  The
local product is done by the ORACLE here (this is a test; the product path is GPU only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spblas_reference_b200 import generators as G
from spblas_reference_b200.sharded import (ShardedSpMV, balanced_nnz_blocks, equal_row_blocks,
                                           plan_exchange)


def test_partitioners():
    assert equal_row_blocks(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert equal_row_blocks(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    rp = torch.tensor([0, 1, 2, 10, 11, 12, 20, 21, 22])
    blocks = balanced_nnz_blocks(rp, 3)
    assert blocks[0][0] == 0 and blocks[-1][1] == 8
    assert all(b[1] == c[0] for b, c in zip(blocks, blocks[1:]))
    nnz = [int(rp[e] - rp[b]) for b, e in blocks]
    assert max(nnz) <= 12                                    # ~22/3 with one 8-entry row each
    assert balanced_nnz_blocks(rp, 1) == [(0, 8)]


def test_exchange_plan_modes():
    blocks = equal_row_blocks(16, 4)
    banded = [(0, 5), (3, 9), (7, 13), (11, 16)]             # each block +- 1 row
    for r in range(4):
        p = plan_exchange(blocks, banded, r)
        assert p.mode == "halo"
        assert all(abs(peer - r) == 1 for peer, _, _ in p.recvs)
        assert p.recv_elems == (1 if r in (0, 3) else 2)
    # what rank r sends to q is exactly what q receives from r
    plans = [plan_exchange(blocks, banded, r) for r in range(4)]
    for r in range(4):
        for peer, b, e in plans[r].sends:
            assert (r, b, e) in plans[peer].recvs
    dense = [(0, 16)] * 4
    assert all(plan_exchange(blocks, dense, r).mode == "allgather" for r in range(4))
    assert plan_exchange([(0, 16)], [(0, 16)], 0).mode == "none"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, kind, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    try:
        if kind == "poisson":
            g = 24
            v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, "cpu")
            blocks = equal_row_blocks(shape[0], world)
        else:
            v, rp, ci, shape = G.rmat_csr(9, 8, seed=5, dtype=torch.float64, device="cpu")
            blocks = balanced_nnz_blocks(rp, world)
        n = shape[1]
        r0, r1 = blocks[rank]
        rp64 = rp.to(torch.int64)
        k0, k1 = int(rp64[r0]), int(rp64[r1])
        lv, lci = v[k0:k1].numpy(), ci[k0:k1].numpy()
        lrp = (rp64[r0:r1 + 1] - k0).to(torch.int32).numpy()
        cols = (int(lci.min()), int(lci.max()) + 1) if len(lci) else (0, 0)

        def local(x, y):
            y.copy_(torch.from_numpy(O.spmv("csr", (r1 - r0, n), lrp, lci, lv, x.numpy(),
                                            alpha_a=0.125)))

        op = ShardedSpMV(n, blocks, cols, local, torch.float64, "cpu")
        x0 = G.dense_uniform((n,), 3, torch.float64, "cpu")
        op.set_x(x0)
        ref = x0.numpy().copy()
        vv, rr, cc = v.numpy(), rp.numpy(), ci.numpy()
        ok = True
        for it in range(4):
            op.step()
            ref = O.spmv("csr", shape, rr, cc, vv, ref, alpha_a=0.125)
            got = op.x_current.numpy()
            if op.plan.mode == "halo":
                lo, hi = cols                                   # only the referenced window is kept current
                ok &= np.array_equal(got[lo:hi], ref[lo:hi])
            else:
                ok &= np.array_equal(got, ref)
            ok &= np.array_equal(op.y_block.numpy(), ref[r0:r1])
        out[rank] = (op.plan.mode, bool(ok), op.plan.recv_elems)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("kind,mode", [("poisson", "halo"), ("rmat", "allgather")])
def test_sharded_iteration(kind, mode, world):
    """world 3 and 4: a middle rank has TWO neighbours (the case that once scattered every tile
    on the GPU path), and 576 rows do not split evenly into R-MAT's nnz-balanced blocks."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), kind, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        got_mode, ok, recv = out[rank]
        assert got_mode == mode and ok
        if kind == "poisson":                                   # one grid line per neighbour
            assert recv == (24 if rank in (0, world - 1) else 48)
