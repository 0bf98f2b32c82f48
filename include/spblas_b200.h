/*
 * spblas_b200.h — C ABI of the B200 (sm_100a) backend for the sparse-times-dense
 * hot path of SparseBLAS/spblas-reference.
 *
 * This is the drop-in boundary: the C++ backend headers under
 * include/spblas/vendor/b200/ (selected with -DSPBLAS_ENABLE_B200, i.e. CMake
 * -DENABLE_B200=ON) decode the reference's views and call ONLY the functions
 * declared here.  Everything behind this header is hand-written CUDA compiled
 * for sm_100a (libspblas_b200.so); there is no cuSPARSE, no CPU fallback.
 *
 * All pointers named `d_*` are DEVICE pointers (the same contract the
 * reference's GPU backends have: test/gtest/device/spmv_test.cpp:27-34 puts
 * raw device pointers in csr_view / std::span).  Scalars (`alpha`) are HOST
 * pointers to one element of the value type.
 *
 * Every entry point returns an int status (0 = success) and never throws.
 * The C++ headers turn statuses into the reference's exception types
 * (include/spblas/vendor/cusparse/exception.hpp:13-21,
 *  include/spblas/algorithms/multiply_impl.hpp:37-41,70-75).
 *
 * Reference interface each entry point replaces:
 *   spblas_b200_plan_create/destroy   <- __cusparse::operation_state_t / spmv_state_t
 *                                        (vendor/cusparse/operation_state_t.hpp:10-37,
 *                                         vendor/cusparse/detail/spmv_state_t.hpp:11-52)
 *   spblas_b200_inspect               <- multiply_inspect (algorithms/multiply.hpp:9-13,29-33;
 *                                        no-op in algorithms/multiply_impl.hpp:19-29,105-116;
 *                                        real work only in vendor/onemkl_sycl/spmv_impl.hpp:34-60)
 *   spblas_b200_spmv                  <- multiply(info, a, x, y) SpMV
 *                                        (algorithms/multiply_impl.hpp:33-62,
 *                                         vendor/cusparse/spmv_impl.hpp:19-90)
 *   spblas_b200_spmm                  <- multiply(info, a, B, C) SpMM
 *                                        (algorithms/multiply_impl.hpp:66-101,
 *                                         vendor/onemkl_sycl/spmm_impl.hpp:88-125)
 *   spblas_b200_plan_query            <- (new) exposes the inspect-phase metadata so the
 *                                        structure tests can compare it bit-exactly
 */
#ifndef SPBLAS_B200_H
#define SPBLAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SPBLAS_B200_API
#else
#define SPBLAS_B200_API __attribute__((visibility("default")))
#endif

#define SPBLAS_B200_VERSION 100

/* ---- status codes ------------------------------------------------------ */
enum {
  SPBLAS_B200_SUCCESS = 0,
  SPBLAS_B200_INVALID_ARGUMENT = 1, /* bad enum / null pointer / negative size      */
  SPBLAS_B200_SHAPE_MISMATCH = 2,   /* -> std::invalid_argument in the C++ header   */
  SPBLAS_B200_NOT_SUPPORTED = 3,    /* type combination not instantiated            */
  SPBLAS_B200_ALLOC_FAILED = 4,     /* -> std::bad_alloc                            */
  SPBLAS_B200_CUDA_ERROR = 5,       /* -> std::runtime_error("CUDA encountered ..") */
  SPBLAS_B200_NOT_INSPECTED = 6,    /* execute called on a plan with no structure   */
  SPBLAS_B200_INVALID_STRUCTURE = 7 /* rowptr not monotone / offsets out of range   */
};

/* ---- enums -------------------------------------------------------------- */
enum { SPBLAS_B200_CSR = 0, SPBLAS_B200_CSC = 1 };
/* index / offset element types (csr_view<T, I, O>: views/csr_view.hpp:12) */
enum { SPBLAS_B200_I32 = 0, SPBLAS_B200_I64 = 1 };
/* value types (vendor/cusparse/types.hpp:17-20 allows fp + int32; complex is rejected) */
enum { SPBLAS_B200_F32 = 0, SPBLAS_B200_F64 = 1, SPBLAS_B200_S32 = 2 };

/* inspect flags */
enum {
  SPBLAS_B200_INSPECT_DEFAULT = 0,
  /* Skip the row-length histogram (partition table and validation of the offsets array
     only).  Used by the no-`info` multiply(a, x, y) overloads. */
  SPBLAS_B200_INSPECT_LIGHT = 1
};

/* plan_query selectors.  All integer outputs are int64_t. */
enum {
  SPBLAS_B200_Q_NUM_TILES = 0,      /* int64[1]: merge-path tiles of the SpMV partition        */
  SPBLAS_B200_Q_TILE_ITEMS = 1,     /* int64[1]: merge items (row ends + nonzeros) per tile    */
  SPBLAS_B200_Q_TILE_STARTS = 2,    /* int64[2*(num_tiles+1)]: (row, nnz) start of every tile  */
  SPBLAS_B200_Q_ROWLEN_HIST = 3,    /* int64[SPBLAS_B200_HIST_BINS]: log2 row-length histogram */
  SPBLAS_B200_Q_MAX_ROW_LEN = 4,    /* int64[1]                                                */
  SPBLAS_B200_Q_EMPTY_ROWS = 5,     /* int64[1]                                                */
  SPBLAS_B200_Q_SPMV_VARIANT = 6,   /* int64[1]: kernel variant chosen for SpMV */
  SPBLAS_B200_Q_LAST_LAUNCHES = 7,  /* int64[1]: kernels launched by the last execute call     */
  SPBLAS_B200_Q_TOTAL_LAUNCHES = 8, /* int64[1]: kernels launched through this plan so far     */
  SPBLAS_B200_Q_CSR_ROWPTR = 9,     /* offset_t[rows+1]: effective CSR rowptr (CSC: transpose) */
  SPBLAS_B200_Q_CSR_COLIND = 10,    /* index_t[nnz]:    effective CSR colind                   */
  SPBLAS_B200_Q_CSR_PERM = 11,      /* offset_t[nnz]:   value gather permutation (CSC only)    */
  SPBLAS_B200_Q_NUM_SEGMENTS = 12,  /* int64[1]: SpMM row segments (0 = rows are not split)    */
  SPBLAS_B200_Q_SEGMENTS = 13,      /* int64[3*num_segments]: (row, nnz_begin, nnz_end)        */
  SPBLAS_B200_Q_SPMM_VARIANT = 14,  /* int64[1]: kernel variant of the last SpMM               */
  SPBLAS_B200_Q_TILE_UNIFORM = 15,  /* int32[num_tiles]: common length L (1..8) of a tile's
                                       complete rows after the first, 0 if they differ          */
  SPBLAS_B200_Q_BARRIER_EPOCH = 16, /* int64[1]: fused-exchange steps signalled so far         */
  SPBLAS_B200_Q_BARRIER_TIMEOUT = 17, /* int64[1]: 1 if a fused barrier gave up waiting for a peer */
  SPBLAS_B200_Q_TRSV_LEVELS = 18,   /* int64[1]: level sets of the inspected triangular solve      */
  SPBLAS_B200_Q_TRSV_SWEEPS = 19,   /* int64[1]: relaxation sweeps the level analysis took          */
  SPBLAS_B200_Q_HUB_COUNT = 20,     /* int64[1]: hub columns held in shared memory (0: none / not built) */
  SPBLAS_B200_Q_HUB_REFS = 21,      /* int64[1]: stored entries that reference a hub column         */
  SPBLAS_B200_Q_HUB_COLS = 22,      /* int32[hub_count]: the hub columns, ascending                 */
  SPBLAS_B200_Q_HUB_COLIND = 23,    /* int32[nnz]: the plan's re-encoded colind (hub number s -> ~s) */
};

/* most destinations / peers of a fused exchange (one NVSwitch domain: 8 GPUs) */
#define SPBLAS_B200_MAX_PEERS 8

/* Row-length histogram: bin 0 = empty rows, bin b >= 1 = rows with
   2^(b-1) <= len < 2^b.  (Row length = rowptr[i+1]-rowptr[i], the definition the
   reference's row() uses: backend/view_customizations.hpp:52-53.) */
#define SPBLAS_B200_HIST_BINS 40

typedef struct spblas_b200_plan spblas_b200_plan;

/* ---- plan lifetime ------------------------------------------------------ */

/* Create an empty plan bound to `cuda_stream` (a cudaStream_t; NULL = legacy
   default stream, which is what the reference's GPU tests rely on:
   test/gtest/device/spmv_test.cpp:34-36 calls multiply then copies back). */
SPBLAS_B200_API int spblas_b200_plan_create(spblas_b200_plan** plan,
                                            void* cuda_stream);
SPBLAS_B200_API void spblas_b200_plan_destroy(spblas_b200_plan* plan);
SPBLAS_B200_API int spblas_b200_plan_set_stream(spblas_b200_plan* plan,
                                                void* cuda_stream);

/* ---- inspect ------------------------------------------------------------ */

/* Analyse the structure of A (m x n, nnz stored entries) on the GPU:
     - validates the offsets array (monotone; ptr[0] may be any base >= 0, as a
       row-block shard of a larger matrix has; ptr[last]-ptr[0] must equal nnz),
     - row-length histogram, max row length, empty-row count,
     - merge-path partition table over (row ends + nonzeros), and per tile whether
       its complete rows all have one length (stencils: no row-end lookups at execute),
     - CSC: builds the row-major (CSR) image of A by a stable counting sort:
       rowptr/colind of the transpose-of-the-storage plus a value permutation,
     - SpMM (k_hint > 1): row segments for rows longer than the segment limit.
   format   SPBLAS_B200_CSR: d_ptr = rowptr[m+1], d_ind = colind[nnz]
            SPBLAS_B200_CSC: d_ptr = colptr[n+1], d_ind = rowind[nnz]
   The plan keeps d_ptr/d_ind (not owned); they must stay alive and structurally
   unchanged until the plan is destroyed or re-inspected.  Values may change
   freely between executes. */
SPBLAS_B200_API int spblas_b200_inspect(spblas_b200_plan* plan, int format,
                                        int64_t m, int64_t n, int64_t nnz,
                                        const void* d_ptr, const void* d_ind,
                                        int off_type, int idx_type,
                                        int64_t k_hint, int flags);

/* Optional, for CSC operands (and transposed(csr), which is one): gather the values
   once into the order of the inspected row-major image, so that later executes stream
   them like a CSR matrix's instead of fetching each through the permutation (a
   scattered 4/8-byte read per entry: 4.6x slower on an R-MAT).  The C++ headers call it
   from multiply_inspect when the operand is wrapped in matrix_opt — the reference's
   marker for "the backend may keep optimised state for this matrix"
   (views/matrix_opt_impl.hpp:14-93; vendor/onemkl_sycl/spmv_impl.hpp:46-53 calls
   optimize_gemv under the same condition).  Contract, as with oneMKL's optimize: the
   values must not change until the next inspect / cache_values.  The cache is used
   only by executes that pass the same d_values pointer and value type; d_values = NULL
   drops it; a no-op for CSR plans. */
SPBLAS_B200_API int spblas_b200_plan_cache_values(spblas_b200_plan* plan, int val_type,
                                                  const void* d_values);

/* Optional: shared-memory residency for the most referenced ("hub") columns of x.
   For matrices whose columns are very unevenly popular (power-law graphs) the product
   is bound by the gathers of x that miss L1, not by HBM.  With enable != 0, a plan that
   would run the general (warp-stream) SpMV kernel counts the references per column on
   its first product, keeps the columns referenced at least min_count times (0 = twice
   the SM count), takes the max_cols most referenced of them (0 = 32768 4-byte or 8192 8-byte
   values: a 164 / 131 KB shared-memory carve-out, which leaves L1 what the gathers in
   flight need; at most 49152 / 20480, what the SM's shared memory holds)
   and stores a re-encoded copy of colind (nnz * 4 bytes, owned by the plan); the hub
   kernel then reads x at those columns from shared memory.  Used only if at least 15 %
   of the stored entries reference a hub; results are bit-identical to the plain kernel's.
   max_cols / min_count = -1 leave the current setting unchanged.  The C++ headers call
   set_hub(1, -1, -1) from multiply_inspect when the operand is wrapped in matrix_opt —
   the reference's marker for "the backend may keep optimised state for this matrix"
   (views/matrix_opt_impl.hpp:14-93).  Measured on R-MAT scale 24, fp32 (C4): 1.13 ms ->
   1.03 ms; the analysis costs 10-50 products, once.
   int32 column indices only; plans inspected through multiply_inspect only (the no-info
   overloads never analyse).  Also read from the environment at plan creation:
   SPBLAS_B200_HUB, SPBLAS_B200_HUB_COLS, SPBLAS_B200_HUB_MIN_COUNT.  No reference
   counterpart (closest: oneMKL's optimize_gemv hint, vendor/onemkl_sycl/spmv_impl.hpp:46-58). */
SPBLAS_B200_API int spblas_b200_plan_set_hub(spblas_b200_plan* plan, int enable,
                                             int64_t max_cols, int64_t min_count);

/* ---- execute ------------------------------------------------------------ */

/* y[m] = alpha * A * x[n]   (beta = 0: y is overwritten, stale contents —
   including NaN — are discarded, as multiply_impl.hpp:43-46 zeroes y first).
   Uses the inspected structure.  Enqueued on the plan's stream; no host
   synchronisation, no allocation. */
SPBLAS_B200_API int spblas_b200_spmv(spblas_b200_plan* plan, int val_type,
                                     const void* alpha, const void* d_values,
                                     const void* d_x, void* d_y);

/* The same product with HOST vectors: h_y[m] = alpha * A * h_x[n].  A (values and the
   inspected structure) stays on the device; d_x[n] and d_y[m] are device staging
   buffers the caller provides (nothing is allocated here).  The upload of x, the
   kernels and the download of y are pipelined chunk by chunk over the plan's tiles:
   chunk c is multiplied as soon as the part of x its columns reach has arrived, and
   its rows travel back while later chunks are multiplied — for a banded matrix PCIe
   is busy in both directions for the whole call.  Bit-identical to spblas_b200_spmv.
   Pinned (page-locked) host memory is needed for the copies to be asynchronous.
   The call is complete in the order of the plan's stream: synchronise that stream
   (or the device) before reading h_y.  No reference counterpart: the reference's GPU
   backends take device pointers only (vendor/cusparse/spmv_impl.hpp:50-66). */
SPBLAS_B200_API int spblas_b200_spmv_host(spblas_b200_plan* plan, int val_type,
                                          const void* alpha, const void* d_values,
                                          const void* h_x, void* h_y, void* d_x,
                                          void* d_y);

/* C[m x k] = alpha * A * B[n x k], B and C row-major with leading dimensions
   ldb, ldc (>= k) in elements.  Two kernels (SPBLAS_B200_Q_SPMM_VARIANT: < 1000 one group of
   lanes per row, >= 1000 merge-path warp streams): the streams when a row of B is at least
   256 bytes, and at any width when the inspect phase's row-length histogram
   (SPBLAS_B200_Q_ROWLEN_HIST) shows heavy-tailed rows — at least 10 % of the stored entries
   in rows of 256 entries or more. */
SPBLAS_B200_API int spblas_b200_spmm(spblas_b200_plan* plan, int val_type,
                                     const void* alpha, const void* d_values,
                                     const void* d_B, int64_t ldb, void* d_C,
                                     int64_t ldc, int64_t k);

/* The 4-argument multiply: y[m] = alpha * A * x[n] + beta * d[m]  and
   C[m x k] = alpha * A * B + beta * D  (D row-major, leading dimension ldd >= k).
   Replaces the form sketched in notes/matrices.hpp (multiply(a, b, c, d)) and implemented by
   the reference only for SpGEMM on rocSPARSE, whose convention it follows
   (vendor/rocsparse/multiply_spgemm.hpp:69-118: alpha = scaling factors of a and b,
   beta = scaling factor of d, i.e. multiply(a, x, y, scaled(beta, d)); :276-283: the
   3-argument form is the 4-argument one with beta = 0).  The addend is fused into the single
   store every row of the result gets — no second pass over y.  d / D may alias y / C (the
   solver update y = alpha A x + beta y).  beta is a HOST pointer; beta == 0 is exactly the
   3-argument product: d is not read, so a NaN in it cannot reach the result. */
SPBLAS_B200_API int spblas_b200_spmv_axpby(spblas_b200_plan* plan, int val_type,
                                           const void* alpha, const void* d_values,
                                           const void* d_x, const void* beta,
                                           const void* d_d, void* d_y);
SPBLAS_B200_API int spblas_b200_spmm_axpby(spblas_b200_plan* plan, int val_type,
                                           const void* alpha, const void* d_values,
                                           const void* d_B, int64_t ldb, const void* beta,
                                           const void* d_D, int64_t ldd, void* d_C,
                                           int64_t ldc, int64_t k);

/* One-shot forms used by the overloads that take no operation_info_t
   (vendor/cusparse/spmv_impl.hpp:92-102 creates and destroys a cuSPARSE handle
   per call there).  They run on a thread-local plan whose buffers are reused.  A structure
   is never ASSUMED unchanged: the first call on it runs a LIGHT inspect (partition +
   validation of the offsets array, no histogram; two host synchronisations); a later SpMV
   call with the same pointers, sizes and types reuses that plan after a DEVICE-side check —
   one kernel recomputes a 64-bit fingerprint of the offsets array and compares it with the
   stored one; the product's kernels are launched right behind it and return without
   touching y unless it matches; the host reads the verdict from host-mapped memory (no
   stream synchronisation) and falls back to the light inspect if the array changed in
   place.  Steady state: one extra kernel over the offsets per call.
   spblas_b200_once_release() frees the calling thread's one-shot plan (it is also freed when
   the thread exits while the CUDA runtime is still loaded). */
SPBLAS_B200_API int spblas_b200_spmv_once(
    void* cuda_stream, int format, int64_t m, int64_t n, int64_t nnz,
    const void* d_ptr, const void* d_ind, int off_type, int idx_type,
    int val_type, const void* alpha, const void* d_values, const void* d_x,
    void* d_y);
SPBLAS_B200_API int spblas_b200_spmm_once(
    void* cuda_stream, int format, int64_t m, int64_t n, int64_t nnz,
    const void* d_ptr, const void* d_ind, int off_type, int idx_type,
    int val_type, const void* alpha, const void* d_values, const void* d_B,
    int64_t ldb, void* d_C, int64_t ldc, int64_t k);
/* ... and their 4-argument forms (y = alpha A x + beta d, C = alpha A B + beta D) */
SPBLAS_B200_API int spblas_b200_spmv_axpby_once(
    void* cuda_stream, int format, int64_t m, int64_t n, int64_t nnz,
    const void* d_ptr, const void* d_ind, int off_type, int idx_type,
    int val_type, const void* alpha, const void* d_values, const void* d_x,
    const void* beta, const void* d_d, void* d_y);
SPBLAS_B200_API int spblas_b200_spmm_axpby_once(
    void* cuda_stream, int format, int64_t m, int64_t n, int64_t nnz,
    const void* d_ptr, const void* d_ind, int off_type, int idx_type,
    int val_type, const void* alpha, const void* d_values, const void* d_B,
    int64_t ldb, const void* beta, const void* d_D, int64_t ldd, void* d_C,
    int64_t ldc, int64_t k);
SPBLAS_B200_API void spblas_b200_once_release(void);

/* ---- triangular_solve(a, uplo, diag, b, x): x = inv(tri(A)) b -----------------

   Replaces triangular_solve_inspect / triangular_solve
   (algorithms/triangular_solve.hpp:8-19, algorithms/triangular_solve_impl.hpp:14-107: a
   no-op inspect and a serial substitution).  A is a general square CSR matrix (m x m);
   only the entries of the chosen triangle and — unless unit_diagonal — the diagonal
   entries are used, the other triangle is ignored, exactly as the reference does.
     trsv_inspect: level sets of the dependency graph, rows ordered by level; with an
       explicit diagonal every row must store one (INVALID_STRUCTURE otherwise: the
       reference divides by the previous row's diagonal there).
     trsv: one launch per level, each x_i computed with the reference's operations in
       the reference's order, every one rounded separately: bit-identical results.
       alpha_a / alpha_b (HOST pointers, NULL = absent) are the factors of scaled(alpha, a)
       and scaled(alpha, b), applied per element as the reference's views do.  d_b may
       alias d_x.  f32 and f64 values. */
SPBLAS_B200_API int spblas_b200_trsv_inspect(spblas_b200_plan* plan, int64_t m, int64_t nnz,
                                             const void* d_rowptr, const void* d_colind,
                                             int off_type, int idx_type, int upper,
                                             int unit_diagonal);
SPBLAS_B200_API int spblas_b200_trsv(spblas_b200_plan* plan, int val_type,
                                     const void* alpha_a, const void* alpha_b,
                                     const void* d_values, const void* d_b, void* d_x);

/* ---- transpose(a, b): B = A^T, both CSR ---------------------------------------

   Replaces transpose_inspect / transpose (algorithms/transpose.hpp:8-13,
   algorithms/transpose_impl.hpp:9-60: count, exclusive scan, scatter, serial).  The
   inspect phase sorts the structure once (stable: within a row of B the entries keep A's
   storage order, exactly the reference's scatter order, so B is bit-identical to the
   reference's); the execute phase copies the structure into B's arrays and moves the
   values through the permutation, so re-transposing after a change of values costs one
   gather pass.  The same plan also multiplies: it IS the plan of transposed(a), usable
   with spblas_b200_spmv / _spmm for y = A^T x.
     transpose_inspect: A is m x n with d_rowptr[m+1], d_colind[nnz].
     transpose:         d_t_rowptr[n+1] (zero-based), d_t_colind[nnz], d_t_values[nnz]. */
SPBLAS_B200_API int spblas_b200_transpose_inspect(spblas_b200_plan* plan, int64_t m,
                                                  int64_t n, int64_t nnz,
                                                  const void* d_rowptr,
                                                  const void* d_colind, int off_type,
                                                  int idx_type);
SPBLAS_B200_API int spblas_b200_transpose(spblas_b200_plan* plan, int val_type,
                                          const void* d_values, void* d_t_rowptr,
                                          void* d_t_colind, void* d_t_values);

/* ---- fused exchange for row-block sharded iterations (y -> x) ----------------

   The reference has no multi-GPU path; this is the B200 side of SURVEY.md 8(e).
   A rank multiplies its row block A[r0:r1, :] against its replica of x; in the
   iterated use x <- alpha A x every rank then needs (parts of) the other ranks'
   rows.  Instead of a collective call after the product, the SpMV kernels store
   the rows a peer needs straight into that peer's buffer over NVLink:

   set_scatter: from now on every spblas_b200_spmv on this plan ALSO stores
     y[i], for row_begin[d] <= i < row_end[d] (rows of THIS plan's block), to
     d_dst[d][i] — d_dst[d] is the peer-mapped address at which this block's
     row 0 lives in destination d's next x (peer's base + r0).  With
     multicast != 0 the addresses are NVLS multicast addresses and the stores
     are multimem.st (one store reaches every GPU).  n_dst = 0 switches it off.
     When the ranges hold few rows in total (a halo: at most
     SPBLAS_B200_LATE_PUSH_ROWS, default 32768) and a barrier is set, the rows are
     copied out of y by the carry fix-up kernel's last block instead of being stored
     by the product kernels (the CTA that owned them was a straggler).
   set_barrier: every spblas_b200_spmv on this plan ends with a flag barrier
     among the ranks: d_remote_slots[q] is this rank's slot in peer q's flag
     array (peer-mapped), d_local_slots[q] the slot peer q writes here; flags
     are uint64 step numbers, zero-initialised by the caller.  Every rank must
     issue the same sequence of executes.  n_peers = 0 switches it off.  A rank that
     waits longer than SPBLAS_B200_BARRIER_TIMEOUT_MS (environment at plan creation,
     default 30000) for a peer gives up instead of hanging the GPU: that step's x is
     incomplete, so the NEXT execute on the plan returns SPBLAS_B200_CUDA_ERROR (the flag
     lives in host-mapped memory: no synchronisation is needed to see it) and
     SPBLAS_B200_Q_BARRIER_TIMEOUT reports 1 from then on.
   Both are properties of the plan (operation_info_t state): multiply /
   multiply_execute keep the reference's signature. */
SPBLAS_B200_API int spblas_b200_plan_set_scatter(spblas_b200_plan* plan, int n_dst,
                                                 void* const* d_dst,
                                                 const int64_t* row_begin,
                                                 const int64_t* row_end,
                                                 int multicast);
SPBLAS_B200_API int spblas_b200_plan_set_barrier(spblas_b200_plan* plan, int n_peers,
                                                 void* const* d_remote_slots,
                                                 const void* const* d_local_slots);

/* ---- introspection / errors ---------------------------------------------- */

/* Copies the selected metadata into `out` (HOST memory, `bytes` capacity).
   Synchronises the plan's stream.  Returns INVALID_ARGUMENT when `bytes` is too
   small; *needed (optional) receives the required size. */
SPBLAS_B200_API int spblas_b200_plan_query(spblas_b200_plan* plan, int what,
                                           void* out, size_t bytes,
                                           size_t* needed);

/* Message of the last failure on this plan ("" if none).  Owned by the plan. */
SPBLAS_B200_API const char* spblas_b200_last_error(const spblas_b200_plan* plan);
/* Message of the last failure of a *_once call on this thread. */
SPBLAS_B200_API const char* spblas_b200_last_error_once(void);
SPBLAS_B200_API const char* spblas_b200_status_string(int status);
SPBLAS_B200_API int spblas_b200_version(void);

/* Measurement aid (csrc/probe.cu), not part of the multiply path: streams d_colind[nnz]
   (and d_values[nnz] unless NULL) with 128-bit loads and gathers d_x through it, one
   accumulator per thread written to d_out[148 * ctas_per_sm * 256 at most 303104
   elements].  bench.py times it to report the gather ceiling of a workload's own
   index stream next to the compulsory-bytes roofline.  Arrays must be 16-byte aligned;
   the tail nnz % 4 is ignored. */
SPBLAS_B200_API int spblas_b200_probe_gather(void* cuda_stream, int idx_type,
                                             int val_type, int64_t nnz,
                                             const void* d_colind, const void* d_values,
                                             const void* d_x, void* d_out,
                                             int ctas_per_sm);

/* Debug/tuning knob (also read from the environment variable
   SPBLAS_B200_SPMV_VARIANT at plan creation): force a SpMV kernel variant for
   ncu A/B runs.  -1 = automatic. */
SPBLAS_B200_API int spblas_b200_plan_force_variant(spblas_b200_plan* plan,
                                                   int spmv_variant);

#ifdef __cplusplus
}
#endif
#endif /* SPBLAS_B200_H */
