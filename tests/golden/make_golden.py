"""Generates tests/golden/*.npz by running the REAL reference (oracle/_ref/libspblas_ref.so,
compiled from /root/reference by oracle/Makefile) on the reference's own fixtures:
spblas::generate_csr / generate_csc / generate_dense with seed 0 over util::dims
(reference test/gtest/util.hpp:27-29), exactly what test/gtest/spmv_test.cpp,
test/gtest/spmm_test.cpp and test/gtest/device/spmv_test.cpp feed to spblas::multiply.

Run in the authoring container (needs /root/reference to have built the _ref library):
    python tests/golden/make_golden.py
The outputs are committed; the GPU box has neither /root/reference nor needs _ref.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DIMS = [(1000, 100, 100), (100, 1000, 10000), (40, 40, 1000)]   # util::dims (m, n, nnz)
SPMM_N = [1, 8, 32, 64, 512]                                     # spmm_test.cpp:11
ALPHAS = [-10, 1, 5]                                             # spmv_test.cpp:45


def main():
    for (m, n, nnz) in DIMS:
        out = {}
        for fmt, gen in (("csr", O.ref_generate_csr), ("csc", O.ref_generate_csc)):
            values, ptr, ind = gen(m, n, nnz, 0, np.float32)
            out[f"{fmt}_values"], out[f"{fmt}_ptr"], out[f"{fmt}_ind"] = values, ptr, ind
            x = np.ones(n, dtype=np.float32)                     # spmv_test.cpp:13
            out[f"{fmt}_spmv"] = O.spmv(fmt, (m, n), ptr, ind, values, x, impl="reference")
            for a in ALPHAS:
                out[f"{fmt}_spmv_ascaled_{a}"] = O.spmv(fmt, (m, n), ptr, ind, values, x,
                                                        alpha_a=a, impl="reference")
                out[f"{fmt}_spmv_bscaled_{a}"] = O.spmv(fmt, (m, n), ptr, ind, values, x,
                                                        alpha_x=a, impl="reference")
            for k in SPMM_N:
                B = O.ref_generate_dense(n, k, 0, np.float32)    # spmm_test.cpp:18
                Cm = O.spmm(fmt, (m, n), ptr, ind, values, B, impl="reference")
                # B is reproducible only through the reference's generator: keep it for
                # the small widths, and a float64 checksum row/column for the wide ones
                if k <= 64:
                    out[f"dense_B_{k}"] = B
                    out[f"{fmt}_spmm_{k}"] = Cm
                    out[f"{fmt}_spmm_ascaled_{k}"] = O.spmm(fmt, (m, n), ptr, ind, values, B,
                                                            alpha_a=2.0, impl="reference")
                else:
                    out[f"dense_B_{k}"] = B
                    out[f"{fmt}_spmm_{k}"] = Cm
        # transpose(a, b) on the CSR fixture (test/gtest/transpose_test.cpp:15-34)
        tv, trp, tci = O.transpose((m, n), out["csr_ptr"], out["csr_ind"], out["csr_values"],
                                   impl="reference")
        out["csr_transpose_values"], out["csr_transpose_ptr"], out["csr_transpose_ind"] = tv, trp, tci
        path = os.path.join(HERE, f"dims_{m}_{n}_{nnz}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")

    # the 3x4 probe of SURVEY.md Appendix A (unsorted row with a duplicate column, an
    # empty row, stale NaN in y, an unreferenced Inf in x)
    rp = np.array([0, 3, 3, 4], np.int32)
    ci = np.array([2, 0, 2, 3], np.int32)
    v = np.array([1, 2, 3, 4], np.float32)
    x = np.array([10, 20, 30, 40], np.float32)
    xinf = x.copy()
    xinf[1] = np.inf
    probe = dict(rowptr=rp, colind=ci, values=v, x=x,
                 y=O.spmv("csr", (3, 4), rp, ci, v, x, impl="reference"),
                 y_inf=O.spmv("csr", (3, 4), rp, ci, v, xinf, impl="reference"),
                 y_scaled=O.spmv("csr", (3, 4), rp, ci, v, x, alpha_a=2, alpha_x=3,
                                 impl="reference"),
                 yt=O.spmv("csc", (4, 3), rp, ci, v, np.ones(3, np.float32), impl="reference"),
                 y_s32=O.spmv("csr", (3, 4), rp, ci, v.astype(np.int32), x.astype(np.int32),
                              impl="reference"))
    np.savez_compressed(os.path.join(HERE, "probe_3x4.npz"), **probe)
    print({k: probe[k].tolist() for k in ("y", "y_inf", "y_scaled", "yt", "y_s32")})


if __name__ == "__main__":
    main()
