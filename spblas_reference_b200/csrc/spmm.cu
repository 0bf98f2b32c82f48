// spmm.cu — C[m x k] = alpha * A * B[n x k], B and C row-major.
//
// Replaces the reference's triple loop
// (include/spblas/algorithms/multiply_impl.hpp:78-91: zero C, then
//  C(i,j) += a_ik * B(k,j) for every stored (i,k), j = 0..k-1).
// The reference has no NVIDIA SpMM path; the row-major contract for B and C is the
// one its vendor SpMM backends use (vendor/onemkl_sycl/spmm_impl.hpp:33-39,116-119).
//
// Mapping: a group of LANES lanes owns one row of A and one tile of LANES*VEC
// columns of C.  Lane l of the group keeps VEC accumulators for columns
// [c0, c0+VEC).  The group loads LANES (colind, value) pairs of the row at a time
// with one coalesced streaming load, then broadcasts them one by one with
// warp shuffles; for every pair each lane issues one 128-bit read-only load of
// its slice of row `col` of B, so a B row is read as full 128-byte lines
// (k = 32 fp32: one line, 8 lanes x float4; k = 128 fp32: four lines, 32 lanes).
// Loads of B for a whole batch of pairs are issued before the first FMA.
// C is written once with 128-bit streaming stores (beta = 0).
//
// Rows longer than kSpmmSegment (hubs of power-law matrices) were cut into
// segments by the inspect phase: each segment is handled by its own group, which
// writes a partial row to a workspace; a second kernel adds a row's partials in
// segment order (deterministic; no floating-point atomics).
#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

namespace {

constexpr int kSpmmThreads = 256;

template <typename T, int VEC>
struct BVec {
  T v[VEC];
};

template <typename T, int VEC>
__device__ __forceinline__ BVec<T, VEC> load_b(const T* p) {
  BVec<T, VEC> r;
  if constexpr (VEC == 1) {
    r.v[0] = ld_ro(p);
  } else {
    static_assert(VEC * sizeof(T) == 16, "vector path moves 16 bytes per lane");
    const uint4 w = ld_ro_16(p);
    *reinterpret_cast<uint4*>(&r.v[0]) = w;
  }
  return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void store_c(T* p, const BVec<T, VEC>& r) {
  if constexpr (VEC == 1) {
    p[0] = r.v[0];
  } else {
    st_stream_16(p, *reinterpret_cast<const uint4*>(&r.v[0]));
  }
}

// Accumulate alpha-less products of the nonzeros [kb, ke) of one row into acc.
template <typename T, typename I, typename O, int VEC, int LANES>
__device__ __forceinline__ void
accumulate_row(const I* __restrict__ colind, const T* __restrict__ values,
               const O* __restrict__ perm, const T* __restrict__ B,
               const int64_t ldb, const int64_t kb, const int64_t ke,
               const int64_t c0, const bool active, const int lane,
               const unsigned gmask, T (&acc)[VEC]) {
  constexpr int BATCH = LANES < 8 ? LANES : 8;
  for (int64_t kk = kb; kk < ke; kk += LANES) {
    I myc = I(0);
    T myv = T(0);
    if (kk + lane < ke) {
      myc = ld_stream(colind + kk + lane);
      myv = perm == nullptr ? ld_stream(values + kk + lane)
                            : ld_ro(values + perm[kk + lane]);
    }
    const int cnt = ke - kk < LANES ? int(ke - kk) : LANES;
#pragma unroll
    for (int j0 = 0; j0 < LANES; j0 += BATCH) {
      if (j0 < cnt) {
        BVec<T, VEC> b[BATCH];
        T v[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
          const I c = __shfl_sync(gmask, myc, j0 + j, LANES);
          v[j] = __shfl_sync(gmask, myv, j0 + j, LANES);
          // Slots past the end of the row must not touch B at all: a NaN/Inf in
          // an unreferenced B row may not leak into C (0 * NaN), exactly as the
          // reference never reads it.  They contribute 0 * 0.
          if (active && j0 + j < cnt) {
            b[j] = load_b<T, VEC>(B + int64_t(c) * ldb + c0);
          } else {
#pragma unroll
            for (int u = 0; u < VEC; ++u)
              b[j].v[u] = T(0);
          }
        }
        if (active) {
#pragma unroll
          for (int j = 0; j < BATCH; ++j)
#pragma unroll
            for (int u = 0; u < VEC; ++u)
              acc[u] += v[j] * b[j].v[u];
        }
      }
    }
  }
}

template <typename T, typename I, typename O, int VEC, int LANES>
__global__ void __launch_bounds__(kSpmmThreads)
spmm_row_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                const T* __restrict__ values, const O* __restrict__ perm,
                const T* __restrict__ B, const int64_t ldb, T* __restrict__ C,
                const int64_t ldc, const T alpha, const int64_t rows,
                const int64_t k, const int64_t seg_limit) {
  constexpr int GROUPS = kSpmmThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int64_t row = int64_t(blockIdx.x) * GROUPS + grp;
  if (row >= rows)
    return; // the whole group leaves together
  const int64_t c0 = (int64_t(blockIdx.y) * LANES + lane) * VEC;
  const bool active = c0 < k;
  const unsigned gmask =
      LANES == 32 ? 0xffffffffu
                  : (((1u << LANES) - 1u) << (((threadIdx.x & 31) / LANES) * LANES));

  const int64_t kb = int64_t(rowptr[row]), ke = int64_t(rowptr[row + 1]);
  T acc[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u)
    acc[u] = T(0);
  // rows cut into segments are produced by spmm_segment_kernel + combine
  if (ke - kb <= seg_limit)
    accumulate_row<T, I, O, VEC, LANES>(colind, values, perm, B, ldb, kb, ke, c0,
                                        active, lane, gmask, acc);
  else
    return;
  if (active) {
    BVec<T, VEC> out;
#pragma unroll
    for (int u = 0; u < VEC; ++u)
      out.v[u] = alpha * acc[u];
    store_c<T, VEC>(C + row * ldc + c0, out);
  }
}

template <typename T, typename I, typename O, int VEC, int LANES>
__global__ void __launch_bounds__(kSpmmThreads)
spmm_segment_kernel(const int64_t* __restrict__ segments,
                    const int64_t num_segments, const I* __restrict__ colind,
                    const T* __restrict__ values, const O* __restrict__ perm,
                    const T* __restrict__ B, const int64_t ldb,
                    T* __restrict__ partial, const int64_t k) {
  constexpr int GROUPS = kSpmmThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int64_t s = int64_t(blockIdx.x) * GROUPS + grp;
  if (s >= num_segments)
    return;
  const int64_t c0 = (int64_t(blockIdx.y) * LANES + lane) * VEC;
  const bool active = c0 < k;
  const unsigned gmask =
      LANES == 32 ? 0xffffffffu
                  : (((1u << LANES) - 1u) << (((threadIdx.x & 31) / LANES) * LANES));
  const int64_t kb = segments[3 * s + 1], ke = segments[3 * s + 2];
  T acc[VEC];
#pragma unroll
  for (int u = 0; u < VEC; ++u)
    acc[u] = T(0);
  accumulate_row<T, I, O, VEC, LANES>(colind, values, perm, B, ldb, kb, ke, c0,
                                      active, lane, gmask, acc);
  if (active) {
    // the partial workspace is dense: row s, leading dimension k
#pragma unroll
    for (int u = 0; u < VEC; ++u)
      partial[s * k + c0 + u] = acc[u];
  }
}

// One thread per (first segment of a split row, column): sums the row's partials
// in segment order and writes alpha * sum to C.
template <typename T>
__global__ void __launch_bounds__(256)
spmm_combine_kernel(const int64_t* __restrict__ segments,
                    const int64_t num_segments, const T* __restrict__ partial,
                    T* __restrict__ C, const int64_t ldc, const T alpha,
                    const int64_t k) {
  const int64_t s = blockIdx.x;
  const int64_t row = segments[3 * s];
  if (s > 0 && segments[3 * (s - 1)] == row)
    return; // not the first segment of its row
  for (int64_t c = threadIdx.x; c < k; c += blockDim.x) {
    T sum = T(0);
    for (int64_t q = s; q < num_segments && segments[3 * q] == row; ++q)
      sum += partial[q * k + c];
    C[row * ldc + c] = alpha * sum;
  }
}

template <typename T, typename I, typename O, int VEC, int LANES>
int launch_spmm(spblas_b200_plan* p, const T alpha, const void* values,
                const void* B, int64_t ldb, void* C, int64_t ldc, int64_t k) {
  constexpr int GROUPS = kSpmmThreads / LANES;
  const int64_t rows = p->csr_rows;
  const int64_t col_tiles = (k + int64_t(LANES) * VEC - 1) / (int64_t(LANES) * VEC);
  const int64_t row_blocks = (rows + GROUPS - 1) / GROUPS;
  if (row_blocks > int64_t(0x7fffffff) || col_tiles > 65535)
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "SpMM grid too large");
  const int64_t seg_limit =
      p->num_segments > 0 ? kSpmmSegment : int64_t(0x7fffffffffffffff);
  int launches = 0;
  if (row_blocks > 0) {
    const dim3 grid{unsigned(row_blocks), unsigned(col_tiles), 1u};
    spmm_row_kernel<T, I, O, VEC, LANES><<<grid, kSpmmThreads, 0, p->stream>>>(
        static_cast<const O*>(p->csr_rowptr), static_cast<const I*>(p->csr_colind),
        static_cast<const T*>(values), static_cast<const O*>(p->csr_perm),
        static_cast<const T*>(B), ldb, static_cast<T*>(C), ldc, alpha, rows, k,
        seg_limit);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmm_row_kernel");
    ++launches;
  }
  if (p->num_segments > 0) {
    int rc = reserve(p, p->seg_partial, size_t(p->num_segments) * size_t(k) * sizeof(T));
    if (rc)
      return rc;
    const int64_t seg_blocks = (p->num_segments + GROUPS - 1) / GROUPS;
    const dim3 grid{unsigned(seg_blocks), unsigned(col_tiles), 1u};
    spmm_segment_kernel<T, I, O, VEC, LANES><<<grid, kSpmmThreads, 0, p->stream>>>(
        static_cast<const int64_t*>(p->segments.p), p->num_segments,
        static_cast<const I*>(p->csr_colind), static_cast<const T*>(values),
        static_cast<const O*>(p->csr_perm), static_cast<const T*>(B), ldb,
        static_cast<T*>(p->seg_partial.p), k);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmm_segment_kernel");
    spmm_combine_kernel<T><<<unsigned(p->num_segments), 256, 0, p->stream>>>(
        static_cast<const int64_t*>(p->segments.p), p->num_segments,
        static_cast<const T*>(p->seg_partial.p), static_cast<T*>(C), ldc, alpha, k);
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmm_combine_kernel");
    launches += 2;
  }
  p->last_launches = launches;
  p->total_launches += launches;
  return SPBLAS_B200_SUCCESS;
}

template <typename T, typename I, typename O>
int pick_shape(spblas_b200_plan* p, const void* alpha, const void* values,
               const void* B, int64_t ldb, void* C, int64_t ldc, int64_t k) {
  constexpr int V = 16 / sizeof(T);
  const T a = *static_cast<const T*>(alpha);
  const auto aligned16 = [](const void* q) {
    return (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  };
  const bool vec = (k % V == 0) && (ldb % V == 0) && (ldc % V == 0) &&
                   aligned16(B) && aligned16(C);
  if (vec) {
    const int64_t nv = k / V;
    if (nv <= 2) {
      p->spmm_variant = 100 * V + 2;
      return launch_spmm<T, I, O, V, 2>(p, a, values, B, ldb, C, ldc, k);
    }
    if (nv <= 8) {
      p->spmm_variant = 100 * V + 8;
      return launch_spmm<T, I, O, V, 8>(p, a, values, B, ldb, C, ldc, k);
    }
    p->spmm_variant = 100 * V + 32;
    return launch_spmm<T, I, O, V, 32>(p, a, values, B, ldb, C, ldc, k);
  }
  if (k <= 2) {
    p->spmm_variant = 102;
    return launch_spmm<T, I, O, 1, 2>(p, a, values, B, ldb, C, ldc, k);
  }
  if (k <= 8) {
    p->spmm_variant = 108;
    return launch_spmm<T, I, O, 1, 8>(p, a, values, B, ldb, C, ldc, k);
  }
  p->spmm_variant = 132;
  return launch_spmm<T, I, O, 1, 32>(p, a, values, B, ldb, C, ldc, k);
}

template <typename T>
int dispatch_index(spblas_b200_plan* p, const void* alpha, const void* values,
                   const void* B, int64_t ldb, void* C, int64_t ldc, int64_t k) {
  const bool i64 = p->idx_type == SPBLAS_B200_I64;
  const bool o64 = p->off_type == SPBLAS_B200_I64;
  if (!i64 && !o64)
    return pick_shape<T, int32_t, int32_t>(p, alpha, values, B, ldb, C, ldc, k);
  if (!i64 && o64)
    return pick_shape<T, int32_t, int64_t>(p, alpha, values, B, ldb, C, ldc, k);
  if (i64 && !o64)
    return pick_shape<T, int64_t, int32_t>(p, alpha, values, B, ldb, C, ldc, k);
  return pick_shape<T, int64_t, int64_t>(p, alpha, values, B, ldb, C, ldc, k);
}

} // namespace

int run_spmm(spblas_b200_plan* p, int val_type, const void* alpha,
             const void* values, const void* B, int64_t ldb, void* C,
             int64_t ldc, int64_t k) {
  p->last_launches = 0;
  if (k == 0 || p->csr_rows == 0)
    return SPBLAS_B200_SUCCESS;
  switch (val_type) {
  case SPBLAS_B200_F32:
    return dispatch_index<float>(p, alpha, values, B, ldb, C, ldc, k);
  case SPBLAS_B200_F64:
    return dispatch_index<double>(p, alpha, values, B, ldb, C, ldc, k);
  case SPBLAS_B200_S32:
    return dispatch_index<int32_t>(p, alpha, values, B, ldb, C, ldc, k);
  default:
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "unknown value type");
  }
}

} // namespace b200
