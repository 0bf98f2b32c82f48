"""Experiment: the host-buffer product of C2 with the kernels reading x from, and writing y to,
PINNED HOST memory directly (UVA: a cudaHostAlloc'ed pointer is valid on the device) against
spblas_b200_spmv_host's chunked copy / multiply / copy pipeline.  x crosses PCIe once only if L2
caches system-memory lines (every element of x is gathered by five rows).  One JSON line."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import spblas_reference_b200 as sb
from spblas_reference_b200 import _cabi, generators as G


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda:0")
    v, rp, ci, shape = G.poisson2d_csr(g, torch.float64, dev)
    n = shape[1]
    a = sb.csr_view(v, rp, ci, shape, int(ci.numel()))
    x0 = G.dense_uniform((n,), 1, torch.float64, dev)
    y_dev = torch.empty(n, dtype=torch.float64, device=dev)
    info = sb.multiply_inspect(a, x0, y_dev)
    a_s = sb.scaled(0.125, a)
    sb.multiply_execute(info, a_s, x0, y_dev)
    torch.cuda.synchronize()
    x_host = x0.cpu().pin_memory()
    y_host = torch.empty(n, dtype=torch.float64).pin_memory()
    y_host2 = torch.empty(n, dtype=torch.float64).pin_memory()

    def chunked():
        sb.multiply_execute_host(info, a_s, x_host, y_host)
        torch.cuda.current_stream().synchronize()

    L = _cabi.lib()
    alpha = np.array([0.125], dtype=np.float64)

    from spblas_reference_b200.views import value_type
    _vt = value_type(v)

    def direct(xp, yp):
        st = L.spblas_b200_spmv(info._plan, _vt, alpha.ctypes.data_as(C.c_void_p), v.data_ptr(),
                                C.c_void_p(xp), C.c_void_p(yp))
        if st != 0:
            raise RuntimeError(info._err())
        torch.cuda.current_stream().synchronize()

    def timeit(fn, reps=8, warm=2):
        for _ in range(warm):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    out = {"exp": "zero_copy", "grid": g, "rows": n}
    out["chunked_ms"] = timeit(chunked)
    out["direct_x_and_y_host_ms"] = timeit(lambda: direct(x_host.data_ptr(), y_host2.data_ptr()))
    out["direct_bit_identical"] = bool(torch.equal(y_host, y_host2))
    out["direct_x_host_y_device_ms"] = timeit(lambda: direct(x_host.data_ptr(), y_dev.data_ptr()))
    out["direct_x_device_y_host_ms"] = timeit(lambda: direct(x0.data_ptr(), y_host2.data_ptr()))
    out["device_only_ms"] = timeit(lambda: direct(x0.data_ptr(), y_dev.data_ptr()))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
