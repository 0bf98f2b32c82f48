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

    # triangular_solve on the reference's own fixtures (test/gtest/triangular_solve_test.cpp:
    # generate_csr over util::square_dims, values scaled by 1e-3) with b = 1, 2, 3, ... so that
    # the answer is not trivially zero, both triangles, implicit unit diagonal; and with an
    # explicit diagonal on the same matrices after a diagonal entry was put in front of every row
    SQUARE_DIMS = [(1000, 1000, 100), (100, 100, 100), (40, 40, 1000)]   # util.hpp:31-33
    out = {}
    for (m, n, nnz) in SQUARE_DIMS:
        values, ptr, ind = O.ref_generate_csr(m, n, nnz, 0, np.float32)
        values = (np.float32(1e-3) * values).astype(np.float32)
        b = (1 + np.arange(m) % 7).astype(np.float32)
        key = f"{m}_{nnz}"
        out[f"values_{key}"], out[f"ptr_{key}"], out[f"ind_{key}"], out[f"b_{key}"] = values, ptr, ind, b
        for upper in (0, 1):
            out[f"x_unit_{'upper' if upper else 'lower'}_{key}"] = O.trsv(
                m, ptr, ind, values, b, upper=upper, unit=True, impl="reference")
        lens = np.diff(ptr)
        dptr = (ptr + np.arange(m + 1)).astype(np.int32)
        dind = np.empty(len(ind) + m, np.int32)
        dval = np.empty(len(ind) + m, np.float32)
        for i in range(m):
            dind[dptr[i]], dval[dptr[i]] = i, np.float32(2 + (i % 3))
            dind[dptr[i] + 1:dptr[i + 1]] = ind[ptr[i]:ptr[i + 1]]
            dval[dptr[i] + 1:dptr[i + 1]] = values[ptr[i]:ptr[i + 1]]
        out[f"dvalues_{key}"], out[f"dptr_{key}"], out[f"dind_{key}"] = dval, dptr, dind
        for upper in (0, 1):
            out[f"x_explicit_{'upper' if upper else 'lower'}_{key}"] = O.trsv(
                m, dptr, dind, dval, b, upper=upper, unit=False, impl="reference")
            out[f"x_explicit_scaled_{'upper' if upper else 'lower'}_{key}"] = O.trsv(
                m, dptr, dind, dval, b, upper=upper, unit=False, alpha_b=1.2, impl="reference")
    np.savez_compressed(os.path.join(HERE, "trsv_square_dims.npz"), **out)
    print("trsv_square_dims.npz", os.path.getsize(os.path.join(HERE, "trsv_square_dims.npz")) // 1024, "KiB")

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
