// spmv.cu — y = alpha * A * x for the effective CSR structure of a plan.
//
// Replaces the serial double loop of the reference
// (include/spblas/algorithms/multiply_impl.hpp:43-52: zero y, then
//  y[i] += a_ik * x[k] in storage order) and the cusparseSpMV call of the
// reference's NVIDIA backend (include/spblas/vendor/cusparse/spmv_impl.hpp:80-84).
//
// Kernel: merge-path tiles.  The inspect phase cut the merged sequence
// (row ends ++ nonzeros) into tiles of kSpmvTileItems items; one CTA owns one
// tile, so every CTA streams the same number of bytes whatever the row-length
// distribution (uniform 5-point stencil rows, Poisson(10) rows, R-MAT hubs).
//   phase 1  row-end offsets of the tile -> shared memory (coalesced)
//   phase 2  colind/values streamed with 128-bit no-L1-allocate loads, x gathered
//            through the read-only path, products -> shared memory
//   phase 3  per-row reduction out of shared memory: thread-per-row for short
//            rows (long rows of such a tile are deferred to a warp-per-row
//            worklist), warp-per-row with a shuffle reduction for long-row tiles
//   phase 4  the tile's trailing partial row becomes a carry (row, value); a
//            second tiny kernel adds carries to y in tile order (deterministic,
//            no floating-point atomics).
// y is written exactly once per row by the tile that holds the row's end, so
// beta = 0 semantics (stale y, even NaN, is discarded) hold without a memset.
#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

namespace {

template <typename T>
struct alignas(16) Vec4 {
  T v[4];
};

constexpr int kLongRow = 64; // thread-per-row tiles hand rows longer than this to warps

template <typename T, typename I, typename O>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_merge_tile_kernel(const O* __restrict__ rowptr,
                       const I* __restrict__ colind,
                       const T* __restrict__ values,
                       const O* __restrict__ perm, const T* __restrict__ x,
                       T* __restrict__ y, const T alpha,
                       const int64_t* __restrict__ tile_starts,
                       const int64_t rows, int64_t* __restrict__ carry_row,
                       T* __restrict__ carry_val, const int vec_ok) {
  constexpr int THREADS = kSpmvThreads;
  constexpr int TILE = kSpmvTileItems;
  constexpr int QITER = TILE / 4 / THREADS;
  constexpr int WARPS = THREADS / 32;
  constexpr int MAXLONG = TILE / kLongRow + 1;

  __shared__ __align__(16) T s_prod[TILE + 4];
  __shared__ int s_rowend[TILE];
  __shared__ T s_red[WARPS];
  __shared__ int s_long[MAXLONG];
  __shared__ int s_nlong;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int64_t t = blockIdx.x;
  const int64_t row0 = tile_starts[2 * t], k0 = tile_starts[2 * t + 1];
  const int64_t row1 = tile_starts[2 * t + 2], k1 = tile_starts[2 * t + 3];
  const int nr = int(row1 - row0);
  const int64_t k0a = k0 & ~int64_t(3); // origin of the smem index (keeps quads aligned)
  const int nz_beg = int(k0 - k0a);
  const int nz_end = int(k1 - k0a);

  if (tid == 0)
    s_nlong = 0;

  // ---- phase 1: row ends ----------------------------------------------------
  for (int q = tid; q < nr; q += THREADS)
    s_rowend[q] = int(int64_t(ld_stream(rowptr + row0 + 1 + q)) - k0a);

  // ---- phase 2: products ------------------------------------------------------
  int64_t ka, kb; // [k0,ka) scalar head, [ka,kb) aligned quads, [kb,k1) scalar tail
  if (vec_ok) {
    ka = (k0 + 3) & ~int64_t(3);
    if (ka > k1)
      ka = k1;
    kb = k1 & ~int64_t(3);
    if (kb < ka)
      kb = ka;
  } else {
    ka = k1;
    kb = k1;
  }
  {
    const int nq = int((kb - ka) >> 2);
    Quad<I> c[QITER];
    Quad<T> v[QITER];
    // issue every streaming load of this thread before the first dependent gather
#pragma unroll
    for (int it = 0; it < QITER; ++it) {
      const int q = tid + it * THREADS;
      if (q < nq) {
        const int64_t k = ka + 4 * int64_t(q);
        c[it] = ld_stream_quad(colind + k);
        if (perm == nullptr) {
          v[it] = ld_stream_quad(values + k);
        } else {
          const Quad<O> pi = ld_stream_quad(perm + k);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[it].v[j] = ld_ro(values + pi.v[j]);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < QITER; ++it) {
      const int q = tid + it * THREADS;
      if (q < nq) {
        Vec4<T> p;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          p.v[j] = v[it].v[j] * ld_ro(x + c[it].v[j]);
        const int64_t k = ka + 4 * int64_t(q);
        *reinterpret_cast<Vec4<T>*>(&s_prod[k - k0a]) = p;
      }
    }
  }
  // scalar head and tail (at most 3 elements each when vec_ok)
  for (int64_t k = k0 + tid; k < ka; k += THREADS) {
    const T a = perm == nullptr ? ld_stream(values + k) : ld_ro(values + perm[k]);
    s_prod[k - k0a] = a * ld_ro(x + ld_stream(colind + k));
  }
  for (int64_t k = kb + tid; k < k1; k += THREADS) {
    const T a = perm == nullptr ? ld_stream(values + k) : ld_ro(values + perm[k]);
    s_prod[k - k0a] = a * ld_ro(x + ld_stream(colind + k));
  }
  __syncthreads();

  // ---- phase 3: complete rows -------------------------------------------------
  const int nzt = nz_end - nz_beg;
  if (nr > 0) {
    if (nzt <= nr * 12) {
      // short rows: one thread per row, sequential (storage-order) sum
      for (int q = tid; q < nr; q += THREADS) {
        const int b = q == 0 ? nz_beg : s_rowend[q - 1];
        const int e = s_rowend[q];
        if (e - b > kLongRow) {
          s_long[atomicAdd(&s_nlong, 1)] = q;
        } else {
          T sum = T(0);
          for (int i = b; i < e; ++i)
            sum += s_prod[i];
          y[row0 + q] = alpha * sum;
        }
      }
      __syncthreads();
      const int nlong = s_nlong;
      for (int w = warp; w < nlong; w += WARPS) {
        const int q = s_long[w];
        const int b = q == 0 ? nz_beg : s_rowend[q - 1];
        const int e = s_rowend[q];
        T sum = T(0);
        for (int i = b + lane; i < e; i += 32)
          sum += s_prod[i];
        sum = warp_reduce_sum(sum);
        if (lane == 0)
          y[row0 + q] = alpha * sum;
      }
    } else {
      // long rows: one warp per row, lanes stride over the row
      for (int q = warp; q < nr; q += WARPS) {
        const int b = q == 0 ? nz_beg : s_rowend[q - 1];
        const int e = s_rowend[q];
        T sum = T(0);
        for (int i = b + lane; i < e; i += 32)
          sum += s_prod[i];
        sum = warp_reduce_sum(sum);
        if (lane == 0)
          y[row0 + q] = alpha * sum;
      }
    }
  }

  // ---- phase 4: trailing partial row -> carry ---------------------------------
  const int tb = nr > 0 ? s_rowend[nr - 1] : nz_beg;
  const int tlen = nz_end - tb;
  if (row1 < rows && tlen > 0) {
    if (tlen <= 256) {
      if (warp == 0) {
        T sum = T(0);
        for (int i = tb + lane; i < nz_end; i += 32)
          sum += s_prod[i];
        sum = warp_reduce_sum(sum);
        if (lane == 0) {
          carry_row[t] = row1;
          carry_val[t] = sum;
        }
      }
    } else {
      T sum = T(0);
      for (int i = tb + tid; i < nz_end; i += THREADS)
        sum += s_prod[i];
      sum = warp_reduce_sum(sum);
      if (lane == 0)
        s_red[warp] = sum;
      __syncthreads();
      if (tid == 0) {
        T tot = T(0);
#pragma unroll
        for (int w = 0; w < WARPS; ++w)
          tot += s_red[w];
        carry_row[t] = row1;
        carry_val[t] = tot;
      }
    }
  } else if (tid == 0) {
    carry_row[t] = -1;
  }
}

// Adds the carries of tiles that ended inside a row to that row's y.  A run of
// consecutive tiles carrying into the same row (a row spanning several tiles) is
// summed in tile order by the thread of the run's first tile.
template <typename T>
__global__ void __launch_bounds__(256)
spmv_carry_fixup_kernel(const int64_t* __restrict__ carry_row,
                        const T* __restrict__ carry_val, int64_t num_tiles,
                        T* __restrict__ y, const T alpha) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= num_tiles)
    return;
  const int64_t r = carry_row[t];
  if (r < 0)
    return;
  if (t > 0 && carry_row[t - 1] == r)
    return;
  T sum = carry_val[t];
  for (int64_t j = t + 1; j < num_tiles && carry_row[j] == r; ++j)
    sum += carry_val[j];
  y[r] += alpha * sum;
}

template <typename T, typename I, typename O>
int launch_spmv(spblas_b200_plan* p, const void* alpha, const void* values,
                const void* x, void* y) {
  if (p->num_tiles == 0)
    return SPBLAS_B200_SUCCESS;
  const T a = *static_cast<const T*>(alpha);
  const auto aligned16 = [](const void* q) {
    return (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  };
  const int vec_ok = aligned16(p->csr_colind) && aligned16(values) &&
                     (p->csr_perm == nullptr || aligned16(p->csr_perm));
  if (p->num_tiles > int64_t(0x7fffffff))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "too many tiles for one launch");
  spmv_merge_tile_kernel<T, I, O>
      <<<unsigned(p->num_tiles), kSpmvThreads, 0, p->stream>>>(
          static_cast<const O*>(p->csr_rowptr),
          static_cast<const I*>(p->csr_colind), static_cast<const T*>(values),
          static_cast<const O*>(p->csr_perm), static_cast<const T*>(x),
          static_cast<T*>(y), a, static_cast<const int64_t*>(p->tile_starts.p),
          p->csr_rows, static_cast<int64_t*>(p->carry_row.p),
          static_cast<T*>(p->carry_val.p), vec_ok);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return cuda_fail(p, e, "spmv_merge_tile_kernel");
  const unsigned fgrid = unsigned((p->num_tiles + 255) / 256);
  spmv_carry_fixup_kernel<T><<<fgrid, 256, 0, p->stream>>>(
      static_cast<const int64_t*>(p->carry_row.p),
      static_cast<const T*>(p->carry_val.p), p->num_tiles, static_cast<T*>(y), a);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return cuda_fail(p, e, "spmv_carry_fixup_kernel");
  p->last_launches = 2;
  p->total_launches += 2;
  return SPBLAS_B200_SUCCESS;
}

template <typename T>
int dispatch_index(spblas_b200_plan* p, const void* alpha, const void* values,
                   const void* x, void* y) {
  const bool i64 = p->idx_type == SPBLAS_B200_I64;
  const bool o64 = p->off_type == SPBLAS_B200_I64;
  if (!i64 && !o64)
    return launch_spmv<T, int32_t, int32_t>(p, alpha, values, x, y);
  if (!i64 && o64)
    return launch_spmv<T, int32_t, int64_t>(p, alpha, values, x, y);
  if (i64 && !o64)
    return launch_spmv<T, int64_t, int32_t>(p, alpha, values, x, y);
  return launch_spmv<T, int64_t, int64_t>(p, alpha, values, x, y);
}

} // namespace

int run_spmv(spblas_b200_plan* p, int val_type, const void* alpha,
             const void* values, const void* x, void* y) {
  p->last_launches = 0;
  p->spmv_variant = kVariantMergeTile;
  switch (val_type) {
  case SPBLAS_B200_F32:
    return dispatch_index<float>(p, alpha, values, x, y);
  case SPBLAS_B200_F64:
    return dispatch_index<double>(p, alpha, values, x, y);
  case SPBLAS_B200_S32:
    return dispatch_index<int32_t>(p, alpha, values, x, y);
  default:
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "unknown value type");
  }
}

} // namespace b200
