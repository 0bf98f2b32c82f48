// trsv.cu — x = inv(tri(A)) b for a square CSR matrix A (SpTRSV), level-scheduled.
//
// Replaces the reference's serial substitution
// (include/spblas/algorithms/triangular_solve_impl.hpp:44-94: rows in solve order, per
// row one pass over the stored entries in storage order — entries on the dependency
// side of the diagonal go into dot_product, an entry on the diagonal sets
// diagonal_value, the other side is ignored — then x_i = (b_i - dot) / diagonal_value,
// or b_i - dot with implicit_unit_diagonal_t) and the no-op triangular_solve_inspect
// (:14-41).
//
// Inspect (GPU): level[i] = 1 + max level of the rows x_i depends on (0 if none), by
// monotone relaxation sweeps until nothing changes; rows ordered by level with a
// stable radix sort (cub, inspect only); offsets of the levels; with an explicit
// diagonal, a check that every row stores one (the reference would divide by the
// PREVIOUS row's diagonal there — diagonal_value is declared outside its row loop,
// :59 — which no parallel schedule can or should reproduce: it is an inspect error
// here).
//
// Execute: one launch per level over that level's rows, one thread per row, entries in
// storage order, every operation rounded separately (__fmul_rn / __fadd_rn / ... : no
// FMA contraction), so that each x_i goes through exactly the arithmetic of the
// reference's loop: the result is BIT-IDENTICAL to the reference's (the oracle and the
// real reference are compiled with -ffp-contract=off), not merely close.  Rows of one
// level never depend on each other, and a launch boundary orders the levels: no flags,
// no spinning, nothing that can hang.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <vector>

#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

namespace {

template <typename I, typename O>
__global__ void __launch_bounds__(256)
trsv_relax_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind, int64_t m,
                  int upper, int* __restrict__ level, int* __restrict__ changed,
                  unsigned long long* __restrict__ max_level) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < m; t += stride) {
    // sweep in solve order so that one sweep carries a level as far as it can
    const int64_t i = upper ? m - 1 - t : t;
    int lvl = 0;
    for (O p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int64_t k = int64_t(colind[p]);
      if (upper ? k > i : k < i) {
        const int lk = reinterpret_cast<volatile int*>(level)[k] + 1;
        lvl = lk > lvl ? lk : lvl;
      }
    }
    if (lvl > level[i]) {
      level[i] = lvl;
      *changed = 1;
      atomicMax(max_level, (unsigned long long)lvl);
    }
  }
}

// stats[1] = rows without a stored diagonal, stats[2] = rows with a column index outside
// [0, m); also fills the identity permutation the sort will carry
template <typename I, typename O>
__global__ void __launch_bounds__(256)
trsv_check_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind, int64_t m,
                  int64_t nnz, int* __restrict__ row_ids,
                  unsigned long long* __restrict__ stats) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  // the offsets first: every later kernel walks colind through them, and the buffers of the
  // column structure are sized by the caller's nnz
  const int64_t first = int64_t(rowptr[0]), last = int64_t(rowptr[m]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && (first < 0 || last - first != nnz))
    atomicAdd(&stats[3], 1ull);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += stride) {
    bool diag = false, bad = false;
    const int64_t lo = int64_t(rowptr[i]), hi = int64_t(rowptr[i + 1]);
    if (hi < lo || lo < first || hi > first + nnz) {
      atomicAdd(&stats[3], 1ull); // not monotone / outside the arrays: do not follow it
      row_ids[i] = int(i);
      continue;
    }
    for (O p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int64_t k = int64_t(colind[p]);
      diag = diag || k == i;
      bad = bad || k < 0 || k >= m;
    }
    row_ids[i] = int(i);
    if (!diag)
      atomicAdd(&stats[1], 1ull);
    if (bad)
      atomicAdd(&stats[2], 1ull);
  }
}

// ---- level sets by frontiers (Kahn): O(nnz) in total, one small launch per level --------
// indeg[i] = stored entries of row i on the dependency side of the diagonal; rows with
// none are level 0 and are appended to `order`.
template <typename I, typename O>
__global__ void __launch_bounds__(256)
trsv_indegree_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind, int64_t m,
                     int upper, int* __restrict__ indeg, int* __restrict__ level,
                     int* __restrict__ order, unsigned long long* __restrict__ count) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += stride) {
    int d = 0;
    for (O p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      const int64_t k = int64_t(colind[p]);
      d += (upper ? k > i : k < i) ? 1 : 0;
    }
    indeg[i] = d;
    if (d == 0) {
      level[i] = 0;
      order[atomicAdd(count, 1ull)] = int(i);
    }
  }
}

// One level: every row k of the frontier order[f0, f1) is solved now, so every row i that
// reads x_k (column k of A: t_rowptr / t_colind, the transpose structure) has one
// dependency less; a row whose count reaches zero joins the next frontier, appended to
// `order` right behind this one and gets its level.  (The frontier list is in whatever
// order the atomics produce; the rows are sorted by (level, row) afterwards so that a
// level's threads walk the matrix in ascending row order.)
template <typename I, typename O>
__global__ void __launch_bounds__(256)
trsv_frontier_kernel(const O* __restrict__ t_rowptr, const I* __restrict__ t_colind,
                     int64_t f0, int64_t f1, int upper, int next_level,
                     int* __restrict__ indeg, int* __restrict__ level, int* __restrict__ order,
                     unsigned long long* __restrict__ count) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t t = f0 + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < f1; t += stride) {
    const int64_t k = order[t];
    for (O p = t_rowptr[k]; p < t_rowptr[k + 1]; ++p) {
      const int64_t i = int64_t(t_colind[p]);
      if (upper ? i < k : i > k)
        if (atomicSub(&indeg[i], 1) == 1) {
          level[i] = next_level;
          order[atomicAdd(count, 1ull)] = int(i);
        }
    }
  }
}

// out[l] = first position in the sorted level array whose level is >= l
__global__ void __launch_bounds__(256)
trsv_level_ptr_kernel(const int* __restrict__ sorted_level, int64_t m, int64_t nlevels,
                      int64_t* __restrict__ out) {
  const int64_t l = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (l > nlevels)
    return;
  int64_t lo = 0, hi = m;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (int64_t(sorted_level[mid]) < l)
      lo = mid + 1;
    else
      hi = mid;
  }
  out[l] = lo;
}

// every operation rounded on its own, as the reference's scalar loop compiled without
// contraction does
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// Operands of a solve as the kernels of a captured graph see them: the graph is built
// once per plan and value type, the operands of each call are copied here first.
struct TrsvParams {
  const void* values;
  const void* b;
  void* x;
  double alpha_a, alpha_b; // (float factors are stored widened; exact)
  int has_aa, has_ab;
};

template <typename T, typename I, typename O>
__device__ __forceinline__ void trsv_row(const O* __restrict__ rowptr,
                                         const I* __restrict__ colind,
                                         const T* __restrict__ values, int64_t i, int upper,
                                         int unit, int has_aa, T alpha_a, int has_ab,
                                         T alpha_b, const T* b, T* x) {
  T dot = T(0), diag = T(0);
  for (O p = rowptr[i]; p < rowptr[i + 1]; ++p) {
    const int64_t k = int64_t(colind[p]);
    T a_v = values[p];
    if (has_aa)
      a_v = mul_rn(alpha_a, a_v);
    if (upper ? k > i : k < i)
      dot = add_rn(dot, mul_rn(a_v, x[k])); // x[k]: written by an earlier launch
    else if (k == i)
      diag = a_v; // the last stored diagonal entry wins, as in the reference
  }
  T b_i = b[i];
  if (has_ab)
    b_i = mul_rn(alpha_b, b_i);
  const T num = sub_rn(b_i, dot);
  x[i] = unit ? num : div_rn(num, diag);
}

template <typename T, typename I, typename O>
__global__ void __launch_bounds__(128)
trsv_level_graph_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                        const int* __restrict__ rows, int64_t nrows, int upper, int unit,
                        const TrsvParams* __restrict__ prm) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nrows)
    return;
  trsv_row<T, I, O>(rowptr, colind, static_cast<const T*>(prm->values), rows[t], upper, unit,
                    prm->has_aa, T(prm->alpha_a), prm->has_ab, T(prm->alpha_b),
                    static_cast<const T*>(prm->b), static_cast<T*>(prm->x));
}

template <typename T, typename I, typename O>
__global__ void __launch_bounds__(128)
trsv_level_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                  const T* __restrict__ values, const int* __restrict__ rows, int64_t nrows,
                  int upper, int unit, int has_aa, T alpha_a, int has_ab, T alpha_b,
                  const T* b, T* x) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nrows)
    return;
  trsv_row<T, I, O>(rowptr, colind, values, rows[t], upper, unit, has_aa, alpha_a, has_ab,
                    alpha_b, b, x);
}

// (A persistent variant — ONE launch, rows in level order, per-row ready flags polled with
// ld.acquire — was written, validated bit-identical on a B200 and measured: 281 ms per solve
// on the 4096^2 stencil against 43.7 ms for the graph replay below, because every one of the
// 8191 wavefronts crosses L2 with an acquire/release pair per row.  It was removed;
// profiles/r02_bench_trsv_persistent_vs_graph.json keeps the measurement.)

} // namespace

void release_trsv_graphs(spblas_b200_plan* p) {
  for (int i = 0; i < 2; ++i) {
    if (p->trsv_graph[i])
      cudaGraphExecDestroy(p->trsv_graph[i]);
    p->trsv_graph[i] = nullptr;
  }
}

namespace {

int check(spblas_b200_plan* p, const char* what) {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPBLAS_B200_SUCCESS : cuda_fail(p, e, what);
}

template <typename I, typename O>
int trsv_inspect_typed(spblas_b200_plan* p, int64_t m, const void* d_rowptr,
                       const void* d_colind, int upper, int unit) {
  cudaStream_t s = p->stream;
  const O* rowptr = static_cast<const O*>(d_rowptr);
  const I* colind = static_cast<const I*>(d_colind);
  if (int rc = reserve(p, p->trsv_level, size_t(std::max<int64_t>(m, 1)) * sizeof(int)))
    return rc;
  if (int rc = reserve(p, p->trsv_order, size_t(std::max<int64_t>(m, 1)) * sizeof(int)))
    return rc;
  if (int rc = reserve(p, p->trsv_tmp0, size_t(std::max<int64_t>(m, 1)) * sizeof(int)))
    return rc;
  if (int rc = reserve(p, p->trsv_tmp1, size_t(std::max<int64_t>(m, 1)) * sizeof(int)))
    return rc;
  if (int rc = reserve(p, p->stats, 64 * sizeof(unsigned long long)))
    return rc;
  int* level = static_cast<int*>(p->trsv_level.p);
  int* order = static_cast<int*>(p->trsv_order.p);
  int* row_ids = static_cast<int*>(p->trsv_tmp0.p);
  int* sorted_level = static_cast<int*>(p->trsv_tmp1.p);
  unsigned long long* stats = static_cast<unsigned long long*>(p->stats.p);
  int* changed = reinterpret_cast<int*>(stats + 8);
  p->trsv_level_ptr_h.assign(1, 0);
  p->trsv_levels = 0;
  release_trsv_graphs(p); // they replay the previous structure's levels
  if (m == 0)
    return SPBLAS_B200_SUCCESS;

  B200_CUDA_TRY(p, cudaMemsetAsync(level, 0, size_t(m) * sizeof(int), s));
  B200_CUDA_TRY(p, cudaMemsetAsync(stats, 0, 16 * sizeof(unsigned long long), s));
  const int grid = int(std::min<int64_t>((m + 255) / 256, int64_t(p->num_sms) * 16));
  // structure first: the sweeps below index level[] with the column indices
  trsv_check_kernel<I, O><<<grid, 256, 0, s>>>(rowptr, colind, m, p->trsv_nnz, row_ids, stats);
  if (int rc = check(p, "trsv_check_kernel"))
    return rc;
  unsigned long long h_stats[4] = {0, 0, 0, 0};
  B200_CUDA_TRY(p, cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  if (h_stats[3] > 0)
    return fail(p, SPBLAS_B200_INVALID_STRUCTURE,
                "offsets array is not monotone or does not span nnz entries");
  if (h_stats[2] > 0)
    return fail(p, SPBLAS_B200_INVALID_STRUCTURE, "column index outside the matrix");
  if (!unit && h_stats[1] > 0)
    return fail(p, SPBLAS_B200_INVALID_STRUCTURE,
                "explicit_diagonal: " + std::to_string(h_stats[1]) +
                    " row(s) store no diagonal entry");
  if (!p->trsv_relax_inspect) {
    // Frontier (Kahn) analysis over the column structure: total work O(nnz), one launch
    // and one 8-byte read-back per level.
    if (int rc = build_column_structure(p, m, p->trsv_nnz, d_rowptr, d_colind))
      return rc;
    const O* t_rowptr = static_cast<const O*>(p->csr_rowptr);
    const I* t_colind = static_cast<const I*>(p->csr_colind);
    int* indeg = sorted_level; // (free until the sort below writes its keys there)
    unsigned long long* count = stats + 4;
    trsv_indegree_kernel<I, O><<<grid, 256, 0, s>>>(rowptr, colind, m, upper, indeg, level,
                                                    order, count);
    if (int rc = check(p, "trsv_indegree_kernel"))
      return rc;
    std::vector<int64_t>& lp = p->trsv_level_ptr_h;
    lp.assign(1, 0);
    unsigned long long done = 0;
    for (;;) {
      B200_CUDA_TRY(p, cudaMemcpyAsync(&done, count, sizeof(done), cudaMemcpyDeviceToHost, s));
      B200_CUDA_TRY(p, cudaStreamSynchronize(s));
      const int64_t f0 = lp.back(), f1 = int64_t(done);
      if (f1 == f0)
        break; // no new rows: all levels found (or a dependency cycle, checked below)
      lp.push_back(f1);
      if (f1 == m)
        break;
      const int fgrid = int(std::min<int64_t>((f1 - f0 + 255) / 256, int64_t(p->num_sms) * 16));
      trsv_frontier_kernel<I, O><<<fgrid, 256, 0, s>>>(t_rowptr, t_colind, f0, f1, upper,
                                                       int(lp.size()) - 1, indeg, level, order,
                                                       count);
      if (int rc = check(p, "trsv_frontier_kernel"))
        return rc;
    }
    if (lp.back() != m)
      return fail(p, SPBLAS_B200_INVALID_STRUCTURE, "level analysis did not reach every row");
    p->trsv_sweeps = int64_t(lp.size()) - 1;
    h_stats[0] = (unsigned long long)(lp.size() - 2); // highest level
  } else {
  // relaxation sweeps: a level can only grow, and it is final once a sweep changes nothing
  for (int64_t sweep = 0;; ++sweep) {
    if (sweep > m)
      return fail(p, SPBLAS_B200_INVALID_STRUCTURE, "level analysis did not converge");
    B200_CUDA_TRY(p, cudaMemsetAsync(changed, 0, sizeof(int), s));
    trsv_relax_kernel<I, O><<<grid, 256, 0, s>>>(rowptr, colind, m, upper, level, changed,
                                                 stats);
    if (int rc = check(p, "trsv_relax_kernel"))
      return rc;
    int h_changed = 0;
    B200_CUDA_TRY(p, cudaMemcpyAsync(&h_changed, changed, sizeof(int), cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
    p->trsv_sweeps = sweep + 1;
    if (!h_changed)
      break;
  }
  B200_CUDA_TRY(p, cudaMemcpyAsync(h_stats, stats, sizeof(h_stats[0]), cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  }
  // rows by (level, row): stable sort of the levels carrying the row ids
  const int64_t nlevels = int64_t(h_stats[0]) + 1;
  int end_bit = 1;
  while (end_bit < 31 && (int64_t(1) << end_bit) < nlevels)
    ++end_bit;
  size_t ws_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, ws_bytes, level, sorted_level, row_ids, order,
                                  int(m), 0, end_bit, s);
  if (int rc = reserve(p, p->sort_ws, ws_bytes))
    return rc;
  B200_CUDA_TRY(p, cub::DeviceRadixSort::SortPairs(p->sort_ws.p, ws_bytes, level, sorted_level,
                                                   row_ids, order, int(m), 0, end_bit, s));
  if (int rc = reserve(p, p->trsv_level_ptr, size_t(nlevels + 1) * sizeof(int64_t)))
    return rc;
  trsv_level_ptr_kernel<<<int((nlevels + 1 + 255) / 256), 256, 0, s>>>(
      sorted_level, m, nlevels, static_cast<int64_t*>(p->trsv_level_ptr.p));
  if (int rc = check(p, "trsv_level_ptr_kernel"))
    return rc;
  p->trsv_level_ptr_h.assign(size_t(nlevels + 1), 0);
  B200_CUDA_TRY(p, cudaMemcpyAsync(p->trsv_level_ptr_h.data(), p->trsv_level_ptr.p,
                                   size_t(nlevels + 1) * sizeof(int64_t),
                                   cudaMemcpyDeviceToHost, s));
  B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  p->trsv_levels = nlevels;
  return SPBLAS_B200_SUCCESS;
}

template <typename T, typename I, typename O>
int trsv_solve_typed(spblas_b200_plan* p, const void* alpha_a, const void* alpha_b,
                     const void* values, const void* b, void* x) {
  const T aa = alpha_a ? *static_cast<const T*>(alpha_a) : T(1);
  const T ab = alpha_b ? *static_cast<const T*>(alpha_b) : T(1);
  const int* order = static_cast<const int*>(p->trsv_order.p);
  int launches = 0;
  // Deep level structures (a 4096 x 4096 stencil has 8191 levels) are bound by launch
  // overhead: the level launches are captured ONCE into a CUDA graph whose kernels read
  // the operands of the call from a small parameter block, and every solve replays it.
  const int slot = sizeof(T) == 8 ? 1 : 0;
  if (p->trsv_use_graph && p->trsv_levels >= 16) {
    if (int rc = reserve(p, p->trsv_params, sizeof(TrsvParams)))
      return rc;
    TrsvParams h;
    h.values = values;
    h.b = b;
    h.x = x;
    h.alpha_a = double(aa);
    h.alpha_b = double(ab);
    h.has_aa = alpha_a != nullptr;
    h.has_ab = alpha_b != nullptr;
    // the previous solve's kernels read the parameter block: if that solve ran on another
    // stream (set_stream between solves), this copy must wait for it
    if (p->trsv_done_event)
      B200_CUDA_TRY(p, cudaStreamWaitEvent(p->stream, p->trsv_done_event, 0));
    B200_CUDA_TRY(p, cudaMemcpyAsync(p->trsv_params.p, &h, sizeof(h), cudaMemcpyHostToDevice,
                                     p->stream));
    if (!p->trsv_graph[slot]) {
      // capture on a private stream: the plan's stream may be the legacy default stream,
      // which cannot be captured
      if (!p->trsv_capture_stream)
        B200_CUDA_TRY(p, cudaStreamCreateWithFlags(&p->trsv_capture_stream, cudaStreamNonBlocking));
      cudaStream_t cs = p->trsv_capture_stream;
      B200_CUDA_TRY(p, cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
      for (int64_t l = 0; l < p->trsv_levels; ++l) {
        const int64_t r0 = p->trsv_level_ptr_h[size_t(l)], r1 = p->trsv_level_ptr_h[size_t(l) + 1];
        if (r1 <= r0)
          continue;
        trsv_level_graph_kernel<T, I, O><<<unsigned((r1 - r0 + 127) / 128), 128, 0, cs>>>(
            static_cast<const O*>(p->trsv_rowptr), static_cast<const I*>(p->trsv_colind),
            order + r0, r1 - r0, p->trsv_upper, p->trsv_unit,
            static_cast<const TrsvParams*>(p->trsv_params.p));
      }
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(cs, &graph);
      if (e != cudaSuccess)
        return cuda_fail(p, e, "cudaStreamEndCapture(trsv levels)");
      e = cudaGraphInstantiate(&p->trsv_graph[slot], graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        p->trsv_graph[slot] = nullptr;
        return cuda_fail(p, e, "cudaGraphInstantiate(trsv levels)");
      }
    }
    B200_CUDA_TRY(p, cudaGraphLaunch(p->trsv_graph[slot], p->stream));
    if (!p->trsv_done_event)
      B200_CUDA_TRY(p, cudaEventCreateWithFlags(&p->trsv_done_event, cudaEventDisableTiming));
    B200_CUDA_TRY(p, cudaEventRecord(p->trsv_done_event, p->stream));
    p->last_launches = p->trsv_levels;
    p->total_launches += p->trsv_levels;
    return SPBLAS_B200_SUCCESS;
  }
  for (int64_t l = 0; l < p->trsv_levels; ++l) {
    const int64_t r0 = p->trsv_level_ptr_h[size_t(l)], r1 = p->trsv_level_ptr_h[size_t(l) + 1];
    if (r1 <= r0)
      continue;
    const int64_t nrows = r1 - r0;
    trsv_level_kernel<T, I, O><<<unsigned((nrows + 127) / 128), 128, 0, p->stream>>>(
        static_cast<const O*>(p->trsv_rowptr), static_cast<const I*>(p->trsv_colind),
        static_cast<const T*>(values), order + r0, nrows, p->trsv_upper, p->trsv_unit,
        alpha_a != nullptr, aa, alpha_b != nullptr, ab, static_cast<const T*>(b),
        static_cast<T*>(x));
    ++launches;
  }
  if (int rc = check(p, "trsv_level_kernel"))
    return rc;
  p->last_launches = launches;
  p->total_launches += launches;
  return SPBLAS_B200_SUCCESS;
}

} // namespace

int trsv_inspect(spblas_b200_plan* p, int64_t m, const void* d_rowptr, const void* d_colind,
                 int upper, int unit) {
  const bool i64 = p->idx_type == SPBLAS_B200_I64, o64 = p->off_type == SPBLAS_B200_I64;
  if (!i64 && !o64)
    return trsv_inspect_typed<int32_t, int32_t>(p, m, d_rowptr, d_colind, upper, unit);
  if (!i64 && o64)
    return trsv_inspect_typed<int32_t, int64_t>(p, m, d_rowptr, d_colind, upper, unit);
  if (i64 && !o64)
    return trsv_inspect_typed<int64_t, int32_t>(p, m, d_rowptr, d_colind, upper, unit);
  return trsv_inspect_typed<int64_t, int64_t>(p, m, d_rowptr, d_colind, upper, unit);
}

template <typename T>
static int trsv_dispatch(spblas_b200_plan* p, const void* alpha_a, const void* alpha_b,
                         const void* values, const void* b, void* x) {
  const bool i64 = p->idx_type == SPBLAS_B200_I64, o64 = p->off_type == SPBLAS_B200_I64;
  if (!i64 && !o64)
    return trsv_solve_typed<T, int32_t, int32_t>(p, alpha_a, alpha_b, values, b, x);
  if (!i64 && o64)
    return trsv_solve_typed<T, int32_t, int64_t>(p, alpha_a, alpha_b, values, b, x);
  if (i64 && !o64)
    return trsv_solve_typed<T, int64_t, int32_t>(p, alpha_a, alpha_b, values, b, x);
  return trsv_solve_typed<T, int64_t, int64_t>(p, alpha_a, alpha_b, values, b, x);
}

int trsv_solve(spblas_b200_plan* p, int val_type, const void* alpha_a, const void* alpha_b,
               const void* values, const void* b, void* x) {
  p->last_launches = 0;
  if (val_type == SPBLAS_B200_F32)
    return trsv_dispatch<float>(p, alpha_a, alpha_b, values, b, x);
  if (val_type == SPBLAS_B200_F64)
    return trsv_dispatch<double>(p, alpha_a, alpha_b, values, b, x);
  return fail(p, SPBLAS_B200_NOT_SUPPORTED, "triangular_solve needs f32 or f64 values");
}

} // namespace b200
