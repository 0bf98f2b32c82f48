"""A/B of the warp-stream kernel (variant 2) and the hub-stream kernel (variant 3) on an
R-MAT matrix: python scripts/hub_ab.py [scale] [fp32|fp64] [cap,cap,...].  One JSON line per
run; cap 0 = the plain walk (variant 2), otherwise the hub table's size in columns."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dtype = torch.float64 if len(sys.argv) > 2 and sys.argv[2] == "fp64" else torch.float32
# (The "g" runs in profiles/r02_hub_ab_l1_bypass.jsonl were L1-bypassing gathers, since made the
# hub kernel's only mode and removed from the plain walk; the "p" runs in
# profiles/r01_hub_ab_rmat.jsonl an index prefetch, since removed: no gain.)
caps = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "49152"]
dev = torch.device("cuda:0")
v, rp, ci, shape = G.rmat_csr(scale, 16, seed=24, dtype=dtype, device=dev)
m, n = shape
nnz = int(ci.numel())
a = sb.csr_view(v, rp, ci, shape, nnz)
x = G.dense_uniform((n,), 5, dtype, dev)
y_plain = None
for spec in caps:
    cap = int(spec)
    y = torch.empty(m, dtype=dtype, device=dev)
    info = sb.multiply_inspect(a, x, y)
    if cap > 0:
        info.set_hub(True, cap, 0)
    info.force_spmv_variant(3 if cap > 0 else 2)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    sb.multiply_execute(info, a, x, y)          # first execute: builds the tables
    t1.record()
    torch.cuda.synchronize()
    first_ms = t0.elapsed_time(t1)
    for _ in range(3):
        sb.multiply_execute(info, a, x, y)
    reps = 30
    t0.record()
    for _ in range(reps):
        sb.multiply_execute(info, a, x, y)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    if cap == 0:
        y_plain = y
    print(json.dumps({"scale": scale, "dtype": str(dtype), "cap": cap, "variant": info.spmv_variant,
                      "ms": round(ms, 4), "first_execute_ms": round(first_ms, 3), "nnz": nnz,
                      "hub_count": info.hub_count,
                      "hub_ref_share": round(info.hub_refs / max(nnz, 1), 4),
                      "gflops": round(2.0 * nnz / ms / 1e6, 1),
                      "max_abs_diff_vs_plain": None if y_plain is None else
                      (y - y_plain).abs().max().item()}), flush=True)
    info.close()
