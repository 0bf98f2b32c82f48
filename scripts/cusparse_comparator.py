"""Same-box comparator for bench.py: NVIDIA cuSPARSE called the way the reference's NVIDIA
backend calls it (include/spblas/vendor/cusparse/spmv_impl.hpp:57-84: generic API,
cusparseSpMV with CUSPARSE_SPMV_ALG_DEFAULT, alpha, beta = 0), via ctypes on the toolkit's
libcusparse.  NOT part of the product — nothing under spblas_reference_b200/ imports this; it
only gives the bench line a "what the reference's GPU backend would do on this box" number.
The comparator is favoured: its workspace is allocated once and only cusparseSpMV / SpMM is
timed, whereas the reference's wrapper also calls bufferSize + cudaMalloc + cudaFree per call
(and creates/destroys a handle per call in the no-info overload, spmv_impl.hpp:99-102).
SpMM has no NVIDIA path in the reference at all; cusparseSpMM (ALG_DEFAULT and CSR_ALG2,
row-major B/C) is measured as the nearest vendor equivalent."""
import ctypes as C
import glob
import os

import torch

_IDX = {torch.int32: 2, torch.int64: 3}            # CUSPARSE_INDEX_32I / 64I
_VAL = {torch.float32: 0, torch.float64: 1}        # CUDA_R_32F / CUDA_R_64F
_ORDER_ROW = 2


def _load():
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cusparse",
                                   "lib", "libcusparse.so*"))
    cands += glob.glob("/usr/local/cuda/lib64/libcusparse.so*")
    for p in cands:
        try:
            return C.CDLL(p)
        except OSError:
            continue
    raise OSError("libcusparse not found")


class CuSparse:
    def __init__(self):
        self.L = _load()
        self.h = C.c_void_p()
        self._chk(self.L.cusparseCreate(C.byref(self.h)), "cusparseCreate")
        self._chk(self.L.cusparseSetStream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "cusparseSetStream")
        self.keep = []

    @staticmethod
    def _chk(st, what):
        if st != 0:
            raise RuntimeError(f"{what} failed with cusparseStatus {st}")

    def _csr(self, m, n, rowptr, colind, values):
        d = C.c_void_p()
        self._chk(self.L.cusparseCreateCsr(
            C.byref(d), C.c_int64(m), C.c_int64(n), C.c_int64(colind.numel()),
            C.c_void_p(rowptr.data_ptr()), C.c_void_p(colind.data_ptr()),
            C.c_void_p(values.data_ptr()), _IDX[rowptr.dtype], _IDX[colind.dtype], 0,
            _VAL[values.dtype]), "cusparseCreateCsr")
        return d

    def spmv(self, m, n, rowptr, colind, values, x, y, alpha=1.0, alg=0, transpose=False):
        """Returns a zero-argument callable that enqueues y = alpha A x (or alpha A^T x:
        CUSPARSE_OPERATION_TRANSPOSE, what get_transpose.hpp:19-29 selects for a CSC base)."""
        op = 1 if transpose else 0
        T = C.c_float if values.dtype == torch.float32 else C.c_double
        a, b = T(alpha), T(0)
        A = self._csr(m, n, rowptr, colind, values)
        X, Y = C.c_void_p(), C.c_void_p()
        self._chk(self.L.cusparseCreateDnVec(C.byref(X), C.c_int64(x.numel()), C.c_void_p(x.data_ptr()),
                                             _VAL[x.dtype]), "cusparseCreateDnVec")
        self._chk(self.L.cusparseCreateDnVec(C.byref(Y), C.c_int64(y.numel()), C.c_void_p(y.data_ptr()),
                                             _VAL[y.dtype]), "cusparseCreateDnVec")
        size = C.c_size_t(0)
        self._chk(self.L.cusparseSpMV_bufferSize(self.h, op, C.byref(a), A, X, C.byref(b), Y,
                                                 _VAL[values.dtype], alg, C.byref(size)),
                  "cusparseSpMV_bufferSize")
        buf = torch.empty(max(size.value, 16), dtype=torch.uint8, device=values.device)
        self.keep += [a, b, buf]
        L, h, ct = self.L, self.h, _VAL[values.dtype]
        bp = C.c_void_p(buf.data_ptr())

        def run():
            st = L.cusparseSpMV(h, op, C.byref(a), A, X, C.byref(b), Y, ct, alg, bp)
            if st != 0:
                raise RuntimeError(f"cusparseSpMV failed with cusparseStatus {st}")
        return run

    def spmm(self, m, n, k, rowptr, colind, values, B, Cm, alpha=1.0, alg=0):
        T = C.c_float if values.dtype == torch.float32 else C.c_double
        a, b = T(alpha), T(0)
        A = self._csr(m, n, rowptr, colind, values)
        Bd, Cd = C.c_void_p(), C.c_void_p()
        self._chk(self.L.cusparseCreateDnMat(C.byref(Bd), C.c_int64(n), C.c_int64(k), C.c_int64(k),
                                             C.c_void_p(B.data_ptr()), _VAL[B.dtype], _ORDER_ROW),
                  "cusparseCreateDnMat")
        self._chk(self.L.cusparseCreateDnMat(C.byref(Cd), C.c_int64(m), C.c_int64(k), C.c_int64(k),
                                             C.c_void_p(Cm.data_ptr()), _VAL[Cm.dtype], _ORDER_ROW),
                  "cusparseCreateDnMat")
        size = C.c_size_t(0)
        self._chk(self.L.cusparseSpMM_bufferSize(self.h, 0, 0, C.byref(a), A, Bd, C.byref(b), Cd,
                                                 _VAL[values.dtype], alg, C.byref(size)),
                  "cusparseSpMM_bufferSize")
        buf = torch.empty(max(size.value, 16), dtype=torch.uint8, device=values.device)
        self.keep += [a, b, buf]
        L, h, ct = self.L, self.h, _VAL[values.dtype]
        bp = C.c_void_p(buf.data_ptr())
        if alg != 0:
            L.cusparseSpMM_preprocess(h, 0, 0, C.byref(a), A, Bd, C.byref(b), Cd, ct, alg, bp)

        def run():
            st = L.cusparseSpMM(h, 0, 0, C.byref(a), A, Bd, C.byref(b), Cd, ct, alg, bp)
            if st != 0:
                raise RuntimeError(f"cusparseSpMM failed with cusparseStatus {st}")
        return run


def time_ms(run, steps, warmup=3):
    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
