"""The alternating loop of the reference's notes (notes/spmv.hpp:12-22) with ONE operation_info_t:
multiply_inspect(info, a, x, y); multiply_inspect(info, transposed(a), y, x); then
multiply_execute(info, a, x, y) / multiply_execute(info, transposed(a), y, x) in turn.  Times a
pair of products with the per-structure plans, and — for comparison — with an info that is made to
forget the other structure before every execute (what the single-plan state did: a re-inspect per
execute).  R-MAT scale 22, fp32, with and without matrix_opt.  One JSON line per case.
(Written at the very end of round 2: its one GPU run lost the first line to a `tail` and stopped at the
matrix_opt case on a misuse of transposed(), fixed here; no number from it is quoted anywhere.)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G


def main():
    M = sys.modules["spblas_reference_b200.multiply"]
    dev = torch.device("cuda:0")
    v, rp, ci, shape = G.rmat_csr(22, 16, seed=24, dtype=torch.float32, device=dev)
    v = v * (1.0 / 64)
    n = shape[0]
    for opt in (False, True):
        a_plain = sb.csr_view(v, rp, ci, tuple(shape), int(ci.numel()))
        a = sb.matrix_opt(a_plain) if opt else a_plain
        # (transposed() takes a plain view, like algorithms/transposed.hpp:7-21)
        at = sb.matrix_opt(sb.transposed(a_plain)) if opt else sb.transposed(a_plain)
        x = G.dense_uniform((n,), 3, torch.float32, dev)
        y = torch.empty(n, dtype=torch.float32, device=dev)
        info = sb.operation_info_t()
        sb.multiply_inspect(info, a, x, y)
        sb.multiply_inspect(info, at, y, x)

        def pair():
            sb.multiply_execute(info, a, x, y)
            sb.multiply_execute(info, at, y, x)

        def forget_parked():
            for _, plan, _ in info._parked:
                M._destroy_plan(plan)
            info._parked = []

        def pair_single_plan():                      # every execute meets an unknown structure
            forget_parked()
            sb.multiply_execute(info, a, x, y)
            forget_parked()
            sb.multiply_execute(info, at, y, x)

        out = {"exp": "alternating", "matrix": "rmat22 fp32", "matrix_opt": opt}
        for name, fn, reps in (("per_structure_plans", pair, 20), ("single_plan", pair_single_plan, 4)):
            x.copy_(G.dense_uniform((n,), 3, torch.float32, dev))
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            out[name + "_ms_per_pair"] = (time.perf_counter() - t0) / reps * 1e3
        print(json.dumps(out), flush=True)
        info.close()


if __name__ == "__main__":
    main()
