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

// Entries of A held per lane for one chunk of a row: a group of LANES lanes keeps
// LANES * kEntries(LANES) (colind, value) pairs in registers.
template <int LANES>
struct ChunkShape {
  static constexpr int E = LANES >= 16 ? 1 : (LANES == 8 ? 2 : 4);
  static constexpr int CHUNK = LANES * E;
};

// coalesced load of the chunk [kk, min(kk + CHUNK, ke)) of a row: entry e of the chunk
// lives in slot e / LANES of lane e % LANES
template <typename T, typename I, typename O, int LANES>
__device__ __forceinline__ void
load_chunk(const I* __restrict__ colind, const T* __restrict__ values,
           const O* __restrict__ perm, int64_t kk, int64_t ke, int lane,
           I (&c)[ChunkShape<LANES>::E], T (&v)[ChunkShape<LANES>::E]) {
#pragma unroll
  for (int sl = 0; sl < ChunkShape<LANES>::E; ++sl) {
    const int64_t idx = kk + sl * LANES + lane;
    c[sl] = I(0);
    v[sl] = T(0);
    if (idx < ke) {
      c[sl] = ld_stream(colind + idx);
      v[sl] = perm == nullptr ? ld_stream(values + idx) : ld_ro(values + perm[idx]);
    }
  }
}

// acc += sum over the first `cnt` entries of the chunk of value * B[col, c0 : c0+VEC]
template <typename T, typename I, int VEC, int LANES>
__device__ __forceinline__ void
process_chunk(const I (&c)[ChunkShape<LANES>::E], const T (&v)[ChunkShape<LANES>::E],
              int cnt, const T* __restrict__ B, int64_t ldb, int64_t c0, bool active,
              unsigned gmask, T (&acc)[VEC]) {
  constexpr int CHUNK = ChunkShape<LANES>::CHUNK;
  constexpr int BATCH = CHUNK < 8 ? CHUNK : 8;
#pragma unroll
  for (int j0 = 0; j0 < CHUNK; j0 += BATCH) {
    if (j0 < cnt) { // group-uniform
      BVec<T, VEC> b[BATCH];
      T a[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        const int e = j0 + j;
        const I col = __shfl_sync(gmask, c[e / LANES], e % LANES, LANES);
        a[j] = __shfl_sync(gmask, v[e / LANES], e % LANES, LANES);
        // Slots past the end of the row must not touch B at all: a NaN/Inf in an
        // unreferenced B row may not leak into C (0 * NaN), exactly as the reference
        // never reads it.  They contribute 0 * 0.
        if (active && e < cnt) {
          b[j] = load_b<T, VEC>(B + int64_t(col) * ldb + c0);
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
            acc[u] += a[j] * b[j].v[u];
      }
    }
  }
}

// Persistent, software-pipelined: every group walks rows row0, row0 + stride, ...  While
// the B rows of the current A row are being gathered, the loads for the FUTURE are
// already in flight — the row offsets of the row two steps ahead and the first chunk of
// (colind, value) pairs of the next row — so that a row costs one round of memory
// latency (the B gather) instead of three (offsets -> pairs -> B).
template <typename T, typename I, typename O, int VEC, int LANES, int MINB>
__global__ void __launch_bounds__(kSpmmThreads, MINB)
spmm_row_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                const T* __restrict__ values, const O* __restrict__ perm,
                const T* __restrict__ B, const int64_t ldb, T* __restrict__ C,
                const int64_t ldc, const T alpha, const int64_t rows,
                const int64_t k, const int64_t seg_limit) {
  constexpr int GROUPS = kSpmmThreads / LANES;
  constexpr int E = ChunkShape<LANES>::E;
  constexpr int CHUNK = ChunkShape<LANES>::CHUNK;
  const int lane = threadIdx.x % LANES;
  const int grp = threadIdx.x / LANES;
  const int64_t stride = int64_t(gridDim.x) * GROUPS;
  int64_t row = int64_t(blockIdx.x) * GROUPS + grp;
  if (row >= rows)
    return; // the whole group leaves together
  const int64_t c0 = (int64_t(blockIdx.y) * LANES + lane) * VEC;
  const bool active = c0 < k;
  const unsigned gmask =
      LANES == 32 ? 0xffffffffu
                  : (((1u << LANES) - 1u) << (((threadIdx.x & 31) / LANES) * LANES));

  int64_t kb = int64_t(rowptr[row]), ke = int64_t(rowptr[row + 1]);
  int64_t kb1 = 0, ke1 = 0;
  if (row + stride < rows) {
    kb1 = int64_t(rowptr[row + stride]);
    ke1 = int64_t(rowptr[row + stride + 1]);
  }
  I pc[E];
  T pv[E];
  load_chunk<T, I, O, LANES>(colind, values, perm, kb, ke, lane, pc, pv);

  for (; row < rows; row += stride) {
    // ---- loads for the future ----------------------------------------------------------
    int64_t kb2 = 0, ke2 = 0;
    if (row + 2 * stride < rows) {
      kb2 = int64_t(rowptr[row + 2 * stride]);
      ke2 = int64_t(rowptr[row + 2 * stride + 1]);
    }
    I nc[E];
    T nv[E];
    load_chunk<T, I, O, LANES>(colind, values, perm, kb1, ke1, lane, nc, nv);

    // ---- the current row ---------------------------------------------------------------
    // rows cut into segments are produced by spmm_segment_kernel + combine
    if (ke - kb <= seg_limit) {
      T acc[VEC];
#pragma unroll
      for (int u = 0; u < VEC; ++u)
        acc[u] = T(0);
      int64_t rem = ke - kb;
      process_chunk<T, I, VEC, LANES>(pc, pv, rem < CHUNK ? int(rem) : CHUNK, B, ldb, c0,
                                      active, gmask, acc);
      for (int64_t kk = kb + CHUNK; kk < ke; kk += CHUNK) {
        I tc[E];
        T tv[E];
        load_chunk<T, I, O, LANES>(colind, values, perm, kk, ke, lane, tc, tv);
        rem = ke - kk;
        process_chunk<T, I, VEC, LANES>(tc, tv, rem < CHUNK ? int(rem) : CHUNK, B, ldb, c0,
                                        active, gmask, acc);
      }
      if (active) {
        BVec<T, VEC> out;
#pragma unroll
        for (int u = 0; u < VEC; ++u)
          out.v[u] = alpha * acc[u];
        store_c<T, VEC>(C + row * ldc + c0, out);
      }
    }
    // ---- rotate the pipeline -------------------------------------------------------------
    kb = kb1;
    ke = ke1;
    kb1 = kb2;
    ke1 = ke2;
#pragma unroll
    for (int sl = 0; sl < E; ++sl) {
      pc[sl] = nc[sl];
      pv[sl] = nv[sl];
    }
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
  for (int64_t kk = kb; kk < ke; kk += ChunkShape<LANES>::CHUNK) {
    I tc[ChunkShape<LANES>::E];
    T tv[ChunkShape<LANES>::E];
    load_chunk<T, I, O, LANES>(colind, values, perm, kk, ke, lane, tc, tv);
    const int64_t rem = ke - kk;
    process_chunk<T, I, VEC, LANES>(
        tc, tv, rem < ChunkShape<LANES>::CHUNK ? int(rem) : ChunkShape<LANES>::CHUNK, B, ldb,
        c0, active, gmask, acc);
  }
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
    // persistent: as many CTAs as are resident at once, each group walking rows
    auto go = [&](auto kern) -> cudaError_t {
      int per_sm = 0;
      cudaError_t eo =
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSpmmThreads, 0);
      if (eo != cudaSuccess || per_sm < 1)
        per_sm = 2;
      const int64_t resident = int64_t(p->num_sms) * per_sm;
      const int64_t gx = row_blocks < resident ? row_blocks : resident;
      const dim3 grid{unsigned(gx), unsigned(col_tiles), 1u};
      kern<<<grid, kSpmmThreads, 0, p->stream>>>(
          static_cast<const O*>(p->csr_rowptr), static_cast<const I*>(p->csr_colind),
          static_cast<const T*>(values), static_cast<const O*>(p->csr_perm),
          static_cast<const T*>(B), ldb, static_cast<T*>(C), ldc, alpha, rows, k, seg_limit);
      return cudaGetLastError();
    };
    // three resident CTAs per SM (<= 80 registers) measured best on C3: two lose
    // gathers in flight, four spill
    cudaError_t e = go(spmm_row_kernel<T, I, O, VEC, LANES, 3>);
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
