// inspect.cu — the GPU inspect phase (multiply_inspect).
//
// The reference's CPU multiply_inspect is a no-op
// (include/spblas/algorithms/multiply_impl.hpp:19-29,105-116); this backend uses
// the phase to analyse the sparsity structure once:
//   1. row-length histogram (log2 bins), max row length, empty rows, and a
//      monotonicity check of the offsets array;
//   2. merge-path partition table: the (row, nnz) coordinate at which every
//      fixed-size tile of the merged sequence (row ends ++ nonzeros) starts;
//   3. CSC input: a row-major (CSR) image of the matrix built by a stable sort of
//      the row indices, plus the permutation that gathers `values` at execute
//      time (values may change between inspect and execute);
//   4. SpMM: the list of row segments for rows longer than kSpmmSegment.
// The CPU restatement these are compared against bit-exactly is oracle/spblas_oracle.c.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

namespace {

constexpr int kStatsWords = SPBLAS_B200_HIST_BINS + 5; // hist, max, flags, nseg, nsplit, uniform tiles

// ---------------------------------------------------------------------------
// 1. row-length histogram
// ---------------------------------------------------------------------------
template <typename O>
__global__ void __launch_bounds__(256)
rowlen_hist_kernel(const O* __restrict__ rowptr, int64_t rows,
                   unsigned long long* __restrict__ stats) {
  __shared__ unsigned int s_hist[SPBLAS_B200_HIST_BINS];
  __shared__ unsigned long long s_max;
  __shared__ unsigned int s_bad;
  if (threadIdx.x < SPBLAS_B200_HIST_BINS)
    s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    s_max = 0;
    s_bad = 0;
  }
  __syncthreads();

  long long local_max = 0;
  bool bad = false;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t rows_padded = (rows + 31) / 32 * 32; // keep warps converged for match
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
       i < rows_padded; i += stride) {
    int bin = -1;
    if (i < rows) {
      const long long a = (long long)rowptr[i];
      const long long b = (long long)rowptr[i + 1];
      const long long len = b - a;
      if (len < 0) {
        bad = true;
        bin = 0;
      } else {
        bin = len == 0 ? 0 : 64 - __clzll(len);
        if (bin > SPBLAS_B200_HIST_BINS - 1)
          bin = SPBLAS_B200_HIST_BINS - 1;
        local_max = len > local_max ? len : local_max;
      }
    }
    // warp-aggregated shared-memory histogram: one atomic per distinct bin
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && (__ffs(peers) - 1) == int(threadIdx.x & 31))
      atomicAdd(&s_hist[bin], __popc(peers));
  }
  for (int off = 16; off > 0; off >>= 1) {
    long long o = __shfl_down_sync(0xffffffffu, local_max, off);
    local_max = o > local_max ? o : local_max;
  }
  if ((threadIdx.x & 31) == 0)
    atomicMax(&s_max, (unsigned long long)local_max);
  if (bad)
    atomicOr(&s_bad, 1u);
  __syncthreads();
  if (threadIdx.x < SPBLAS_B200_HIST_BINS && s_hist[threadIdx.x] != 0)
    atomicAdd(&stats[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
  if (threadIdx.x == 0) {
    atomicMax(&stats[SPBLAS_B200_HIST_BINS], s_max);
    if (s_bad)
      atomicOr(&stats[SPBLAS_B200_HIST_BINS + 1], 1ull);
  }
}

// ---------------------------------------------------------------------------
// 1b. fingerprint of the offsets array (the no-info overloads' structure cache)
// ---------------------------------------------------------------------------
// multiply(a, x, y) without an operation_info_t may not assume that the structure is the
// one it saw last time — but re-deriving the partition costs two host synchronisations per
// call.  So the one-shot plan keeps the last structure's plan together with a 64-bit
// fingerprint of its offsets array (order-independent sum of mixed (index, offset) pairs;
// everything the plan derives — partition, uniform-tile table, warp streams — is a function of
// the offsets alone).  A later call with the same pointers and sizes launches THIS kernel
// first: it recomputes the fingerprint, compares it ON THE DEVICE with the stored one, and
// opens the gate the SpMV kernels of the same call check (gate == seq: run; else: return
// without touching y).  The verdict also goes to host-mapped memory, where the host reads it
// without a stream synchronisation.
//   store mode (light inspect): also validates (monotone, base >= 0, last - first == nnz).
struct HashState {            // device memory, zero-initialised
  unsigned long long acc;     // running sum of the launch
  unsigned int blocks_done;
  unsigned int gate;          // sequence number of the last verified call, or 0
  unsigned long long stored;  // fingerprint kept by the light inspect
  unsigned int bad;           // store mode: 1 = offsets array is malformed
  unsigned int pad;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename O>
__global__ void __launch_bounds__(256)
offsets_fingerprint_kernel(const O* __restrict__ rowptr, const int64_t rows, const int64_t nnz,
                           HashState* __restrict__ st, const int compare,
                           const unsigned int seq, unsigned long long* host_status) {
  __shared__ unsigned long long s_part[8];
  __shared__ unsigned int s_bad;
  if (threadIdx.x == 0)
    s_bad = 0;
  __syncthreads();
  unsigned long long h = 0;
  bool bad = false;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  // four independent loads in flight per thread: the pass is latency-bound otherwise (13 us
  // for a 4 MB array with one load per thread and iteration, ncu)
  for (int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 <= rows; i0 += 4 * stride) {
    long long v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      v[u] = i <= rows ? (long long)ld_stream(rowptr + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i > rows)
        break;
      h += mix64((unsigned long long)v[u] * 0xD6E8FEB86659FD93ull + (unsigned long long)i);
      if (!compare) {
        if (i < rows && (long long)rowptr[i + 1] < v[u])
          bad = true;
        if (i == 0 && (v[u] < 0 || (long long)rowptr[rows] - v[u] != (long long)nnz))
          bad = true;
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1)
    h += __shfl_down_sync(0xffffffffu, h, off);
  if ((threadIdx.x & 31) == 0)
    s_part[threadIdx.x >> 5] = h;
  if (bad)
    atomicOr(&s_bad, 1u);
  __syncthreads();
  if (threadIdx.x != 0)
    return;
  unsigned long long tot = 0;
  for (int w = 0; w < int(blockDim.x >> 5); ++w)
    tot += s_part[w];
  atomicAdd(&st->acc, tot);
  if (s_bad)
    atomicOr(&st->bad, 1u);
  __threadfence();
  if (atomicAdd(&st->blocks_done, 1u) != gridDim.x - 1)
    return;
  // the last block: the launch's fingerprint is complete
  __threadfence();
  const unsigned long long fp = atomicAdd(&st->acc, 0ull);
  st->acc = 0;
  st->blocks_done = 0;
  if (compare) {
    const bool same = fp == st->stored;
    st->gate = same ? seq : 0u;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(host_status) =
        ((unsigned long long)seq << 1) | (same ? 1ull : 0ull);
  } else {
    st->stored = fp;
    st->gate = 0u;
  }
}

// ---------------------------------------------------------------------------
// 2. merge-path partition (Merrill & Garland): list A = row-end offsets
//    rowptr[1..rows] - base, list B = the natural numbers 0..nnz-1.  Tile t
//    starts at diagonal t * tile_items.
// ---------------------------------------------------------------------------
template <typename O>
__global__ void __launch_bounds__(256)
merge_partition_kernel(const O* __restrict__ rowptr, int64_t rows, int64_t nnz,
                       int64_t base, int tile_items, int64_t num_tiles,
                       int64_t* __restrict__ tile_starts) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t > num_tiles)
    return;
  int64_t d = t * int64_t(tile_items);
  const int64_t total = rows + nnz;
  if (d > total)
    d = total;
  int64_t lo = d > nnz ? d - nnz : 0;
  int64_t hi = d < rows ? d : rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t row_end = int64_t(rowptr[mid + 1]) - base;
    if (row_end <= d - 1 - mid)
      lo = mid + 1;
    else
      hi = mid;
  }
  tile_starts[2 * t] = lo;
  tile_starts[2 * t + 1] = base + (d - lo);
}

// Per tile: do all complete rows after the first (rows row0+1 .. row1-1, whose ends
// lie in the tile) have the same length L, 1 <= L <= 8?  Then out[t] = L, else 0.
// One warp per tile.  Stencil-like and fixed-degree matrices are uniform almost
// everywhere; the execute kernel then needs no row-end lookups for the tile.
template <typename O>
__global__ void __launch_bounds__(256)
tile_uniform_kernel(const O* __restrict__ rowptr,
                    const int64_t* __restrict__ tile_starts, int64_t num_tiles,
                    int max_len, int* __restrict__ out,
                    unsigned long long* __restrict__ uniform_count) {
  const int64_t t = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= num_tiles)
    return;
  const int64_t row0 = tile_starts[2 * t], row1 = tile_starts[2 * t + 2];
  int64_t L = -1;
  if (row1 - row0 >= 2)
    L = int64_t(rowptr[row0 + 2]) - int64_t(rowptr[row0 + 1]);
  bool ok = L >= 1 && L <= max_len;
  if (ok) {
    for (int64_t r = row0 + 1 + lane; r < row1; r += 32)
      ok = ok && (int64_t(rowptr[r + 1]) - int64_t(rowptr[r]) == L);
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    out[t] = ok ? int(L) : 0;
    if (ok)
      atomicAdd(uniform_count, 1ull);
  }
}

// ---------------------------------------------------------------------------
// 3. CSC -> row-major image
// ---------------------------------------------------------------------------
// entry e of the CSC storage lives in column j: write j for every e in
// [colptr[j], colptr[j+1]) (a "column of entry" expansion).
template <typename I, typename O>
__global__ void __launch_bounds__(256)
expand_major_kernel(const O* __restrict__ ptr, int64_t majors, int64_t base,
                    I* __restrict__ major_of_entry, O* __restrict__ entry_id,
                    int64_t nnz) {
  // one warp per major index; lanes stride over its entries
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t j = warp; j < majors; j += nwarps) {
    const int64_t b = int64_t(ptr[j]) - base, e = int64_t(ptr[j + 1]) - base;
    for (int64_t p = b + lane; p < e; p += 32) {
      major_of_entry[p] = I(j);
      entry_id[p] = O(p + base);
    }
  }
}

// rowptr of the sorted row keys: out[i] = lower_bound(keys, i)
template <typename I, typename O>
__global__ void __launch_bounds__(256)
rowptr_from_sorted_kernel(const I* __restrict__ keys, int64_t nnz, int64_t rows,
                          O* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > rows)
    return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (int64_t(keys[mid]) < i)
      lo = mid + 1;
    else
      hi = mid;
  }
  out[i] = O(lo);
}

template <typename I, typename O>
__global__ void __launch_bounds__(256)
gather_cols_kernel(const I* __restrict__ major_of_entry,
                   const O* __restrict__ perm, int64_t base, int64_t nnz,
                   I* __restrict__ out) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < nnz)
    out[p] = major_of_entry[int64_t(perm[p]) - base];
}

// ---------------------------------------------------------------------------
// 4. SpMM row segments: rows longer than `seg` are cut into ceil(len/seg) pieces.
//    Pass 1 counts, pass 2 (after a host-side size check) fills in row order.
// ---------------------------------------------------------------------------
template <typename O>
__global__ void __launch_bounds__(256)
count_segments_kernel(const O* __restrict__ rowptr, int64_t rows, int64_t seg,
                      unsigned long long* __restrict__ stats) {
  unsigned long long nseg = 0, nsplit = 0;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += stride) {
    const int64_t len = int64_t(rowptr[i + 1]) - int64_t(rowptr[i]);
    if (len > seg) {
      nseg += (unsigned long long)((len + seg - 1) / seg);
      nsplit += 1;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    nseg += __shfl_down_sync(0xffffffffu, nseg, off);
    nsplit += __shfl_down_sync(0xffffffffu, nsplit, off);
  }
  if ((threadIdx.x & 31) == 0 && nseg) {
    atomicAdd(&stats[SPBLAS_B200_HIST_BINS + 2], nseg);
    atomicAdd(&stats[SPBLAS_B200_HIST_BINS + 3], nsplit);
  }
}

int launch_ok(spblas_b200_plan* p, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return cuda_fail(p, e, what);
  return SPBLAS_B200_SUCCESS;
}

template <typename O>
int launch_fingerprint(spblas_b200_plan* p, const O* rowptr, int64_t rows, int compare,
                       unsigned int seq) {
  if (!p->fp_state.p) {
    if (int rc = reserve(p, p->fp_state, sizeof(HashState)))
      return rc;
    B200_CUDA_TRY(p, cudaMemsetAsync(p->fp_state.p, 0, sizeof(HashState), p->stream));
  }
  if (!compare)
    B200_CUDA_TRY(p, cudaMemsetAsync(static_cast<char*>(p->fp_state.p) + offsetof(HashState, bad),
                                     0, sizeof(unsigned int), p->stream));
  // four offsets per thread and pass; the launch ends with one atomic per CTA on a single word
  const int64_t want = (rows + 1 + 1023) / 1024;
  const int grid = int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(p->num_sms) * 8)));
  offsets_fingerprint_kernel<O><<<grid, 256, 0, p->stream>>>(
      rowptr, rows, p->nnz, static_cast<HashState*>(p->fp_state.p), compare, seq,
      p->fp_status_d);
  return launch_ok(p, "offsets_fingerprint_kernel");
}

template <typename O>
int inspect_rows(spblas_b200_plan* p, const O* rowptr, int64_t rows, int flags) {
  cudaStream_t s = p->stream;
  const bool light_pre = (flags & SPBLAS_B200_INSPECT_LIGHT) != 0;

  // The light inspect (no-info overloads) skips the histogram, not the validation: one pass
  // over the offsets checks them and leaves their fingerprint for the structure cache.
  if (light_pre && rows > 0)
    if (int rc = launch_fingerprint<O>(p, rowptr, rows, 0, 0u))
      return rc;
  // base offset and total: two scalars from the offsets array
  O ends[2] = {0, 0};
  unsigned int malformed = 0;
  if (rows > 0) {
    B200_CUDA_TRY(p, cudaMemcpyAsync(&ends[0], rowptr, sizeof(O),
                                     cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaMemcpyAsync(&ends[1], rowptr + rows, sizeof(O),
                                     cudaMemcpyDeviceToHost, s));
    if (light_pre)
      B200_CUDA_TRY(p, cudaMemcpyAsync(&malformed,
                                       static_cast<char*>(p->fp_state.p) + offsetof(HashState, bad),
                                       sizeof(malformed), cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  }
  p->base = int64_t(ends[0]);
  if (p->base < 0 || int64_t(ends[1]) - p->base != p->nnz)
    return fail(p, SPBLAS_B200_INVALID_STRUCTURE,
                "offsets array does not span nnz entries (ptr[last]-ptr[0] != nnz)");
  if (malformed)
    return fail(p, SPBLAS_B200_INVALID_STRUCTURE,
                "offsets array is not monotonically non-decreasing");

  int rc = reserve(p, p->stats, kStatsWords * sizeof(unsigned long long));
  if (rc)
    return rc;
  auto* d_stats = static_cast<unsigned long long*>(p->stats.p);
  B200_CUDA_TRY(p, cudaMemsetAsync(d_stats, 0,
                                   kStatsWords * sizeof(unsigned long long), s));

  const bool light = (flags & SPBLAS_B200_INSPECT_LIGHT) != 0;
  p->have_hist = false;
  if (!light && rows > 0) {
    const int64_t want = (rows + 255) / 256;
    const int grid = int(std::min<int64_t>(want, int64_t(p->num_sms) * 8));
    rowlen_hist_kernel<O><<<grid, 256, 0, s>>>(rowptr, rows, d_stats);
    if (int e = launch_ok(p, "rowlen_hist_kernel"))
      return e;
    count_segments_kernel<O><<<grid, 256, 0, s>>>(rowptr, rows, kSpmmSegment,
                                                   d_stats);
    if (int e = launch_ok(p, "count_segments_kernel"))
      return e;
  }

  // merge-path partition table
  if (p->tile_items_override > 0)
    p->tile_items = p->tile_items_override;
  else
    p->tile_items = kSpmvTileItems;
  const int64_t total = rows + p->nnz;
  p->num_tiles = (total + p->tile_items - 1) / p->tile_items;
  rc = reserve(p, p->tile_starts, size_t(p->num_tiles + 1) * 2 * sizeof(int64_t));
  if (rc)
    return rc;
  rc = reserve(p, p->carry_row, size_t(p->num_tiles + 1) * sizeof(int64_t));
  if (rc)
    return rc;
  rc = reserve(p, p->carry_val, size_t(p->num_tiles + 1) * 8);
  if (rc)
    return rc;
  {
    const int64_t nthreads = p->num_tiles + 1;
    const int grid = int((nthreads + 255) / 256);
    merge_partition_kernel<O><<<grid, 256, 0, s>>>(
        rowptr, rows, p->nnz, p->base, p->tile_items, p->num_tiles,
        static_cast<int64_t*>(p->tile_starts.p));
    if (int e = launch_ok(p, "merge_partition_kernel"))
      return e;
  }
  rc = reserve(p, p->tile_uniform, size_t(p->num_tiles + 1) * sizeof(int));
  if (rc)
    return rc;
  if (p->num_tiles > 0) {
    const int64_t nthreads = p->num_tiles * 32;
    const int grid = int((nthreads + 255) / 256);
    tile_uniform_kernel<O><<<grid, 256, 0, s>>>(
        rowptr, static_cast<const int64_t*>(p->tile_starts.p), p->num_tiles, 8,
        static_cast<int*>(p->tile_uniform.p), d_stats + SPBLAS_B200_HIST_BINS + 4);
    if (int e = launch_ok(p, "tile_uniform_kernel"))
      return e;
    // how many tiles take the stencil path decides which SpMV kernel runs
    unsigned long long nuni = 0;
    B200_CUDA_TRY(p, cudaMemcpyAsync(&nuni, d_stats + SPBLAS_B200_HIST_BINS + 4,
                                     sizeof(nuni), cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
    p->uniform_tiles = int64_t(nuni);
  } else {
    p->uniform_tiles = 0;
  }

  if (!light && rows > 0) {
    unsigned long long h[kStatsWords];
    B200_CUDA_TRY(p, cudaMemcpyAsync(h, d_stats, sizeof(h),
                                     cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
    if (h[SPBLAS_B200_HIST_BINS + 1] != 0)
      return fail(p, SPBLAS_B200_INVALID_STRUCTURE,
                  "offsets array is not monotonically non-decreasing");
    for (int b = 0; b < SPBLAS_B200_HIST_BINS; ++b)
      p->hist[b] = int64_t(h[b]);
    p->max_row_len = int64_t(h[SPBLAS_B200_HIST_BINS]);
    p->empty_rows = p->hist[0];
    p->have_hist = true;
    p->num_segments = int64_t(h[SPBLAS_B200_HIST_BINS + 2]);
    p->num_split_rows = int64_t(h[SPBLAS_B200_HIST_BINS + 3]);
  } else {
    for (int b = 0; b < SPBLAS_B200_HIST_BINS; ++b)
      p->hist[b] = 0;
    p->max_row_len = 0;
    p->empty_rows = 0;
    p->num_segments = 0;
    p->num_split_rows = 0;
  }
  return SPBLAS_B200_SUCCESS;
}

// minor indices of a CSC operand must lie inside the matrix: the radix sort below looks at
// the low bits only, an index outside would silently corrupt the image
template <typename I>
__global__ void __launch_bounds__(256)
index_range_kernel(const I* __restrict__ ind, int64_t nnz, int64_t bound,
                   unsigned long long* __restrict__ flag) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  bool bad = false;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int64_t v = int64_t(ld_stream(ind + k));
    bad = bad || v < 0 || v >= bound;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0)
    atomicOr(flag, 1ull);
}

template <typename I, typename O>
int build_row_major_image(spblas_b200_plan* p) {
  // CSC storage: ptr = colptr[n+1], ind = rowind[nnz].  Stable radix sort of the
  // row indices (keys) carrying the storage position (values) gives, for every
  // row, its entries in ascending (column, storage) order — the same order in
  // which the reference's column-major scatter (backend/algorithms.hpp:21-29)
  // adds them into c[i].
  cudaStream_t s = p->stream;
  const int64_t nnz = p->nnz, rows = p->m, cols = p->n;
  const O* colptr = static_cast<const O*>(p->user_ptr);
  const I* rowind = static_cast<const I*>(p->user_ind);

  O first = 0;
  if (cols > 0) {
    B200_CUDA_TRY(p, cudaMemcpyAsync(&first, colptr, sizeof(O),
                                     cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  }
  const int64_t base = int64_t(first);

  int rc;
  if ((rc = reserve(p, p->own_rowptr, size_t(rows + 1) * sizeof(O))))
    return rc;
  if ((rc = reserve(p, p->own_colind, size_t(std::max<int64_t>(nnz, 1)) * sizeof(I))))
    return rc;
  if ((rc = reserve(p, p->own_perm, size_t(std::max<int64_t>(nnz, 1)) * sizeof(O))))
    return rc;
  if ((rc = reserve(p, p->sort_tmp0, size_t(std::max<int64_t>(nnz, 1)) * sizeof(I))))
    return rc; // column of entry
  if ((rc = reserve(p, p->sort_tmp1, size_t(std::max<int64_t>(nnz, 1)) * sizeof(O))))
    return rc; // entry id (unsorted)
  if ((rc = reserve(p, p->sort_tmp2, size_t(std::max<int64_t>(nnz, 1)) * sizeof(I))))
    return rc; // sorted keys

  I* col_of_entry = static_cast<I*>(p->sort_tmp0.p);
  O* entry_id = static_cast<O*>(p->sort_tmp1.p);
  I* sorted_keys = static_cast<I*>(p->sort_tmp2.p);
  O* perm = static_cast<O*>(p->own_perm.p);
  O* t_rowptr = static_cast<O*>(p->own_rowptr.p);
  I* t_colind = static_cast<I*>(p->own_colind.p);

  if (nnz > 0) {
    {
      const int64_t warps = std::min<int64_t>(cols, int64_t(p->num_sms) * 64);
      const int grid = int(std::max<int64_t>(1, (warps * 32 + 255) / 256));
      expand_major_kernel<I, O><<<grid, 256, 0, s>>>(colptr, cols, base,
                                                     col_of_entry, entry_id, nnz);
      if (int e = launch_ok(p, "expand_major_kernel"))
        return e;
    }
    if (nnz > int64_t(0x7fffffff))
      return fail(p, SPBLAS_B200_NOT_SUPPORTED,
                  "CSC inspect supports at most 2^31-1 stored entries");
    {
      if ((rc = reserve(p, p->seg_counter, sizeof(unsigned long long))))
        return rc;
      auto* flag = static_cast<unsigned long long*>(p->seg_counter.p);
      B200_CUDA_TRY(p, cudaMemsetAsync(flag, 0, sizeof(unsigned long long), s));
      const int grid = int(std::min<int64_t>((nnz + 255) / 256, int64_t(p->num_sms) * 16));
      index_range_kernel<I><<<grid, 256, 0, s>>>(rowind + base, nnz, rows, flag);
      if (int e = launch_ok(p, "index_range_kernel"))
        return e;
      unsigned long long outside = 0;
      B200_CUDA_TRY(p, cudaMemcpyAsync(&outside, flag, sizeof(outside), cudaMemcpyDeviceToHost, s));
      B200_CUDA_TRY(p, cudaStreamSynchronize(s));
      if (outside)
        return fail(p, SPBLAS_B200_INVALID_STRUCTURE, "index outside the matrix");
    }
    int end_bit = 1;
    while (end_bit < int(sizeof(I) * 8) && (int64_t(1) << end_bit) < rows)
      ++end_bit;
    size_t ws_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, ws_bytes, rowind + base,
                                    sorted_keys, entry_id, perm, int(nnz), 0,
                                    end_bit, s);
    if ((rc = reserve(p, p->sort_ws, ws_bytes)))
      return rc;
    B200_CUDA_TRY(p, cub::DeviceRadixSort::SortPairs(
                         p->sort_ws.p, ws_bytes, rowind + base, sorted_keys,
                         entry_id, perm, int(nnz), 0, end_bit, s));
    {
      const int grid = int((nnz + 255) / 256);
      gather_cols_kernel<I, O><<<grid, 256, 0, s>>>(col_of_entry, perm, base,
                                                    nnz, t_colind);
      if (int e = launch_ok(p, "gather_cols_kernel"))
        return e;
    }
  }
  {
    const int grid = int((rows + 1 + 255) / 256);
    rowptr_from_sorted_kernel<I, O><<<grid, 256, 0, s>>>(sorted_keys, nnz, rows,
                                                         t_rowptr);
    if (int e = launch_ok(p, "rowptr_from_sorted_kernel"))
      return e;
  }
  p->csr_rowptr = t_rowptr;
  p->csr_colind = t_colind;
  p->csr_perm = perm;
  p->csr_rows = rows;
  p->csr_cols = cols;
  return SPBLAS_B200_SUCCESS;
}

template <typename O>
__global__ void __launch_bounds__(256)
fill_segments_kernel(const O* __restrict__ rowptr, int64_t rows, int64_t seg,
                     unsigned long long* __restrict__ cursor,
                     int64_t* __restrict__ segments) {
  // Deterministic order is restored by the host sort below; this kernel only
  // needs every split row to emit its pieces contiguously.
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += stride) {
    const int64_t b = int64_t(rowptr[i]), e = int64_t(rowptr[i + 1]);
    const int64_t len = e - b;
    if (len > seg) {
      const unsigned long long n = (unsigned long long)((len + seg - 1) / seg);
      unsigned long long at = atomicAdd(cursor, n);
      for (unsigned long long q = 0; q < n; ++q) {
        const int64_t sb = b + int64_t(q) * seg;
        segments[3 * (at + q) + 0] = i;
        segments[3 * (at + q) + 1] = sb;
        segments[3 * (at + q) + 2] = sb + seg < e ? sb + seg : e;
      }
    }
  }
}

template <typename O>
int build_segments(spblas_b200_plan* p, const O* rowptr, int64_t rows) {
  if (p->num_segments == 0)
    return SPBLAS_B200_SUCCESS;
  cudaStream_t s = p->stream;
  int rc;
  if ((rc = reserve(p, p->segments, size_t(p->num_segments) * 3 * sizeof(int64_t))))
    return rc;
  if ((rc = reserve(p, p->seg_counter, sizeof(unsigned long long))))
    return rc;
  B200_CUDA_TRY(p, cudaMemsetAsync(p->seg_counter.p, 0, sizeof(unsigned long long), s));
  const int64_t want = (rows + 255) / 256;
  const int grid = int(std::min<int64_t>(want, int64_t(p->num_sms) * 8));
  fill_segments_kernel<O><<<grid, 256, 0, s>>>(
      rowptr, rows, kSpmmSegment,
      static_cast<unsigned long long*>(p->seg_counter.p),
      static_cast<int64_t*>(p->segments.p));
  if (int e = launch_ok(p, "fill_segments_kernel"))
    return e;
  // Split rows are rare (hubs of power-law matrices); order the list by
  // (row, begin) on the host so that the layout — and with it the order in
  // which partial rows are combined — is deterministic.
  std::vector<int64_t> h(size_t(p->num_segments) * 3);
  B200_CUDA_TRY(p, cudaMemcpyAsync(h.data(), p->segments.p,
                                   h.size() * sizeof(int64_t),
                                   cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  struct Seg {
    int64_t r, b, e;
  };
  auto* segs = reinterpret_cast<Seg*>(h.data());
  std::sort(segs, segs + p->num_segments, [](const Seg& a, const Seg& b) {
    return a.r != b.r ? a.r < b.r : a.b < b.b;
  });
  B200_CUDA_TRY(p, cudaMemcpyAsync(p->segments.p, h.data(),
                                   h.size() * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, s));
  B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  return SPBLAS_B200_SUCCESS;
}

template <typename I, typename O>
int inspect_typed(spblas_b200_plan* p, int flags) {
  if (p->format == SPBLAS_B200_CSC) {
    if (int rc = build_row_major_image<I, O>(p))
      return rc;
  } else {
    p->csr_rowptr = p->user_ptr;
    p->csr_colind = p->user_ind;
    p->csr_perm = nullptr;
    p->csr_rows = p->m;
    p->csr_cols = p->n;
  }
  const O* rowptr = static_cast<const O*>(p->csr_rowptr);
  if (int rc = inspect_rows<O>(p, rowptr, p->csr_rows, flags))
    return rc;
  if (int rc = build_segments<O>(p, rowptr, p->csr_rows))
    return rc;
  return SPBLAS_B200_SUCCESS;
}

template <typename O>
int stream_partition_typed(spblas_b200_plan* p, int64_t streams) {
  const int64_t total = p->csr_rows + p->nnz;
  int64_t items = (total + streams - 1) / streams;
  if (items < 1)
    items = 1;
  if (items > int64_t(0x7fffffff))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "SpMM stream too long");
  if (int rc = reserve(p, p->spmm_starts, size_t(streams + 1) * 2 * sizeof(int64_t)))
    return rc;
  if (int rc = reserve(p, p->spmm_carry_row, size_t(streams) * sizeof(int64_t)))
    return rc;
  const int grid = int((streams + 1 + 255) / 256);
  merge_partition_kernel<O><<<grid, 256, 0, p->stream>>>(
      static_cast<const O*>(p->csr_rowptr), p->csr_rows, p->nnz, p->base, int(items), streams,
      static_cast<int64_t*>(p->spmm_starts.p));
  if (int e = launch_ok(p, "merge_partition_kernel (SpMM streams)"))
    return e;
  p->spmm_streams = streams;
  return SPBLAS_B200_SUCCESS;
}

template <typename O>
int ws_partition_typed(spblas_b200_plan* p, int items, int64_t streams) {
  if (int rc = reserve(p, p->ws_starts, size_t(streams + 1) * 2 * sizeof(int64_t)))
    return rc;
  if (int rc = reserve(p, p->ws_carry_row, size_t(streams) * sizeof(int64_t)))
    return rc;
  if (int rc = reserve(p, p->ws_carry_val, size_t(streams) * 8))
    return rc;
  const int grid = int((streams + 1 + 255) / 256);
  merge_partition_kernel<O><<<grid, 256, 0, p->stream>>>(
      static_cast<const O*>(p->csr_rowptr), p->csr_rows, p->nnz, p->base, items, streams,
      static_cast<int64_t*>(p->ws_starts.p));
  if (int e = launch_ok(p, "merge_partition_kernel (SpMV warp streams)"))
    return e;
  p->ws_streams = streams;
  p->ws_items = items;
  return SPBLAS_B200_SUCCESS;
}

} // namespace

// The warp-stream table of spmv_warp_stream_kernel: the merged sequence cut into runs
// of `items` merge items, one run per resident warp when the matrix is small, runs of
// 4096 items dealt round-robin to the warps when it is large.
int build_ws_partition(spblas_b200_plan* p, int64_t resident_warps) {
  const int64_t total = p->csr_rows + p->nnz;
  int64_t items = (total + resident_warps - 1) / resident_warps;
  if (p->ws_items_override > 0)
    items = p->ws_items_override;
  else if (items > 4096)
    items = 4096;
  if (items < 256)
    items = 256;
  const int64_t streams = total > 0 ? (total + items - 1) / items : 0;
  if (streams > int64_t(0x7fffffff))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "too many warp streams");
  if (streams == 0) {
    p->ws_streams = 0;
    return SPBLAS_B200_SUCCESS;
  }
  return p->off_type == SPBLAS_B200_I64 ? ws_partition_typed<int64_t>(p, int(items), streams)
                                        : ws_partition_typed<int32_t>(p, int(items), streams);
}

// The SpMM stream table: the same merge-path cut as the SpMV tiles, but into exactly
// `streams` runs (one per resident warp of spmm_ring_kernel).
int build_stream_partition(spblas_b200_plan* p, int64_t streams) {
  return p->off_type == SPBLAS_B200_I64 ? stream_partition_typed<int64_t>(p, streams)
                                        : stream_partition_typed<int32_t>(p, streams);
}

namespace {

// b_values[q] = a_values[perm[q]]: the value half of transpose(a, b).  perm is streamed,
// the values are gathered (each is read exactly once), the output is streamed.
template <typename W, typename O>
__global__ void __launch_bounds__(256)
gather_values_kernel(const W* __restrict__ values, const O* __restrict__ perm, int64_t nnz,
                     W* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < nnz; q += stride)
    out[q] = ld_ro(values + int64_t(ld_stream(perm + q)));
}

template <typename O>
int transpose_typed(spblas_b200_plan* p, size_t val_bytes, const void* values, void* t_values) {
  const int64_t nnz = p->nnz;
  if (nnz == 0)
    return SPBLAS_B200_SUCCESS;
  const int grid = int(std::min<int64_t>((nnz + 255) / 256, int64_t(p->num_sms) * 16));
  const O* perm = static_cast<const O*>(p->csr_perm);
  if (val_bytes == 8)
    gather_values_kernel<uint64_t, O><<<grid, 256, 0, p->stream>>>(
        static_cast<const uint64_t*>(values), perm, nnz, static_cast<uint64_t*>(t_values));
  else
    gather_values_kernel<uint32_t, O><<<grid, 256, 0, p->stream>>>(
        static_cast<const uint32_t*>(values), perm, nnz, static_cast<uint32_t*>(t_values));
  return launch_ok(p, "gather_values_kernel");
}

} // namespace

// transpose(info, a, b) after transpose_inspect: the plan holds the CSR structure of
// A^T (it was inspected as the CSC matrix A^T over A's own arrays); the structure is
// copied out and the values follow through the permutation.
int run_transpose(spblas_b200_plan* p, int val_type, const void* values, void* t_rowptr,
                  void* t_colind, void* t_values) {
  const size_t so = type_size_idx(p->off_type), si = type_size_idx(p->idx_type);
  B200_CUDA_TRY(p, cudaMemcpyAsync(t_rowptr, p->csr_rowptr, size_t(p->csr_rows + 1) * so,
                                   cudaMemcpyDeviceToDevice, p->stream));
  if (p->nnz > 0)
    B200_CUDA_TRY(p, cudaMemcpyAsync(t_colind, p->csr_colind, size_t(p->nnz) * si,
                                     cudaMemcpyDeviceToDevice, p->stream));
  return p->off_type == SPBLAS_B200_I64
             ? transpose_typed<int64_t>(p, type_size_val(val_type), values, t_values)
             : transpose_typed<int32_t>(p, type_size_val(val_type), values, t_values);
}

// The column structure of a square CSR matrix — rowptr/colind of its transpose — in the
// plan's own_* buffers (the CSC-image builder run on A's arrays read column-major).
// Used by the triangular solve's level analysis: column k lists the rows that read x_k.
int build_column_structure(spblas_b200_plan* p, int64_t m, int64_t nnz, const void* d_rowptr,
                           const void* d_colind) {
  p->format = SPBLAS_B200_CSC;
  p->m = m;
  p->n = m;
  p->nnz = nnz;
  p->user_ptr = d_rowptr;
  p->user_ind = d_colind;
  const bool i64 = p->idx_type == SPBLAS_B200_I64, o64 = p->off_type == SPBLAS_B200_I64;
  if (!i64 && !o64)
    return build_row_major_image<int32_t, int32_t>(p);
  if (!i64 && o64)
    return build_row_major_image<int32_t, int64_t>(p);
  if (i64 && !o64)
    return build_row_major_image<int64_t, int32_t>(p);
  return build_row_major_image<int64_t, int64_t>(p);
}

// out[q] = values[perm[q]] for the whole image (the value half of run_transpose)
int gather_permuted_values(spblas_b200_plan* p, int val_type, const void* values, void* out) {
  return p->off_type == SPBLAS_B200_I64
             ? transpose_typed<int64_t>(p, type_size_val(val_type), values, out)
             : transpose_typed<int32_t>(p, type_size_val(val_type), values, out);
}

// The cached one-shot structure: recompute the fingerprint of the caller's offsets array and
// compare it on the device with the one the light inspect stored (CSR plans only).
int verify_structure(spblas_b200_plan* p, unsigned int seq) {
  if (p->off_type == SPBLAS_B200_I64)
    return launch_fingerprint<int64_t>(p, static_cast<const int64_t*>(p->user_ptr), p->m, 1, seq);
  return launch_fingerprint<int32_t>(p, static_cast<const int32_t*>(p->user_ptr), p->m, 1, seq);
}

const unsigned int* structure_gate(const spblas_b200_plan* p) {
  return p->fp_state.p ? reinterpret_cast<const unsigned int*>(
                             static_cast<const char*>(p->fp_state.p) + offsetof(HashState, gate))
                       : nullptr;
}

int inspect_structure(spblas_b200_plan* p, int flags) {
  const bool i64 = p->idx_type == SPBLAS_B200_I64;
  const bool o64 = p->off_type == SPBLAS_B200_I64;
  int rc;
  p->spmm_streams = 0; // the stream tables belong to the previous structure
  p->ws_streams = -1;
  if (!i64 && !o64)
    rc = inspect_typed<int32_t, int32_t>(p, flags);
  else if (!i64 && o64)
    rc = inspect_typed<int32_t, int64_t>(p, flags);
  else if (i64 && !o64)
    rc = inspect_typed<int64_t, int32_t>(p, flags);
  else
    rc = inspect_typed<int64_t, int64_t>(p, flags);
  return rc;
}

} // namespace b200
