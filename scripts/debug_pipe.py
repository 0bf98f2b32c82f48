"""Debug helper: run the pipelined SpMV on a mid-size Poisson problem and compare with the
oracle (use under compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spblas_reference_b200 as sb
from spblas_reference_b200 import generators as G
from oracle import oracle as O

g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, dev)
a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
x = G.dense_uniform((shape[1],), 1, torch.float64, dev)
y = torch.full((shape[0],), float("nan"), dtype=torch.float64, device=dev)
info = sb.multiply_inspect(a, x, y)
print("tiles", info.num_tiles, "tile_items", info.tile_items)
sb.multiply_execute(info, a, x, y)
torch.cuda.synchronize()
print("variant", info.spmv_variant)
ref = O.spmv("csr", shape, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy(), x.cpu().numpy())
err = np.abs(y.cpu().numpy() - ref)
print("max err", err.max(), "bad rows", int((err > 1e-12).sum()), np.nonzero(err > 1e-12)[0][:10])
