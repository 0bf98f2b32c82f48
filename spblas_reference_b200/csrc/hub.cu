// hub.cu — inspect-phase analysis of COLUMN popularity for spmv_hub_stream_kernel.
//
// The reference's multiply gathers x[j] once per stored entry
// (include/spblas/algorithms/multiply_impl.hpp:48-52); on a GPU every such gather that
// misses L1 costs a request on the SM's port to L2, and that port — not HBM — bounds
// matrices with scattered columns (csrc/probe.cu).  When a few columns
// take a large share of the references (power-law graphs), x at those columns can live
// in shared memory instead.  This file finds them:
//
//   1. count the references to every column (one atomic per stored entry, inspect only);
//   2. keep the columns referenced at least `min_count` times (a hub column costs one
//      load per CTA and launch, so it must be referenced far more often than there are
//      CTAs), order them by (count descending, column ascending) with a stable radix
//      sort, and take the first `cap` — what the kernel's shared memory holds;
//   3. write the plan's own copy of the effective colind in which a reference to hub
//      number s (hubs renumbered in ascending column order) is stored as ~s.
//
// The caller's arrays are never modified; the copy costs nnz * 4 bytes in the plan.
// Only int32 column indices (a negative index marks a hub), which is what every named
// workload uses.  Nothing here runs at execute time except on the first product that
// asks for the hub variant (the capacity depends on the value width).
#include <algorithm>
#include <numeric>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

namespace {

__global__ void __launch_bounds__(256)
hub_count_kernel(const int32_t* __restrict__ colind, int64_t nnz, int64_t cols,
                 int* __restrict__ counts) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int32_t c = ld_stream(colind + k);
    if (c >= 0 && int64_t(c) < cols) // a column outside the matrix is the caller's bug, not a crash here
      atomicAdd(counts + c, 1);
  }
}

struct AtLeast {
  const int* counts;
  int min_count;
  __device__ __forceinline__ bool operator()(const int32_t& col) const {
    return counts[col] >= min_count;
  }
};

__global__ void __launch_bounds__(256)
hub_gather_counts_kernel(const int* __restrict__ counts, const int32_t* __restrict__ cand,
                         int64_t n, int* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = counts[cand[i]];
}

__global__ void __launch_bounds__(256)
hub_slot_kernel(const int32_t* __restrict__ hub_cols, int64_t h, int* __restrict__ slot_of) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < h)
    slot_of[hub_cols[i]] = int(i);
}

// out[pad + k] = ~slot if column colind[k] is a hub, else colind[k]; out[0 .. pad) = 0
__global__ void __launch_bounds__(256)
hub_encode_kernel(const int32_t* __restrict__ colind, int64_t nnz, int64_t cols,
                  const int* __restrict__ slot_of, int pad, int32_t* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nnz + pad; k += stride) {
    int32_t v = 0;
    if (k >= pad) {
      v = ld_stream(colind + (k - pad));
      if (v >= 0 && int64_t(v) < cols) {
        const int s = slot_of[v];
        v = s >= 0 ? ~s : v;
      }
    }
    out[k] = v;
  }
}

int launch_check(spblas_b200_plan* p, const char* what) {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPBLAS_B200_SUCCESS : cuda_fail(p, e, what);
}

struct ToInt64 {
  __device__ __forceinline__ int64_t operator()(const int& v) const { return int64_t(v); }
};

struct Scratch { // freed when the analysis returns, however it returns
  DeviceBuffer counts, cand, cand_counts, sorted_counts, sorted_cols, ws, nsel;
  ~Scratch() {
    for (DeviceBuffer* b : {&counts, &cand, &cand_counts, &sorted_counts, &sorted_cols, &ws, &nsel})
      release(*b);
  }
};

} // namespace

// How many columns of x the hub kernel holds.  The hardware limit is the SM's shared
// memory minus the walk's per-warp slabs (256 entries each); the DEFAULT stops at 163 KB
// of shared memory, because every gather in flight occupies an L1 line and L1 is what
// shared memory leaves of the SM's unified array: on R-MAT scale 24 (fp32) 32768 columns
// (160 KB) run in 1.03 ms, 40960 (192 KB) in 1.08 ms, 49152 (224 KB) in 1.76 ms — slower
// than no table at all (1.13 ms).  With the walk's branch-free load phase (eight gathers in
// flight per lane instead of four) the optimum moved towards a larger L1
// (profiles/r02_hub_sweep_flat.jsonl): fp32 24576 / 32768 columns 0.971 / 0.975 ms, 40960
// 1.128; fp64 (R-MAT scale 24) 8192 columns 1.231 ms, 12288 1.334, 16384 1.706 — the
// 8-byte default is 8192 columns (131 KB of shared memory).  Whole 1024-column steps.
int64_t hub_capacity(const spblas_b200_plan* p, size_t val_bytes, int walk_warps) {
  const int64_t slabs = int64_t(walk_warps) * 256 * int64_t(val_bytes);
  const auto columns = [&](int64_t budget) {
    const int64_t c = (budget - slabs) / int64_t(val_bytes);
    return c > 0 ? (c / 1024) * 1024 : int64_t(0);
  };
  const int64_t hw = columns(int64_t(p->smem_per_sm) - 1024); // per-CTA limit: 1 KB is the system's
  if (p->hub_cap_override > 0)
    return p->hub_cap_override < hw ? p->hub_cap_override : hw;
  return std::min(hw, columns(int64_t(val_bytes == 8 ? 131 : 163) * 1024));
}

// How many columns the GLOBAL-memory table holds (spmv_hubg_stream_kernel): half of L2 —
// the table is what should stay resident beside the streams of A.
int64_t hub_global_capacity(const spblas_b200_plan* p, size_t val_bytes) {
  if (p->hub_cap_override > 0)
    return p->hub_cap_override;
  return (p->l2_bytes / 2) / int64_t(val_bytes);
}

// by_popularity = false: the table of the shared-memory kernel (hubs in ascending column
// order); true: the table of the global-memory kernel (hubs in descending popularity, ties in
// ascending column order — the hottest entries share cache lines, so L1 holds the top of it).
int build_hub_table(spblas_b200_plan* p, int64_t cap, bool by_popularity) {
  p->hub_state = -1;
  p->hub_count = 0;
  p->hub_refs = 0;
  p->hub_cap = cap;
  p->hub_by_popularity = by_popularity;
  if (p->idx_type != SPBLAS_B200_I32)
    return SPBLAS_B200_SUCCESS; // no hubs: the caller falls back to the warp-stream kernel
  const int64_t nnz = p->nnz, cols = p->csr_cols;
  const int pad = int(p->base & 3);
  cudaStream_t s = p->stream;
  const int32_t* colind = static_cast<const int32_t*>(p->csr_colind) + p->base;
  if (int rc = reserve(p, p->hub_colind, size_t(nnz + pad) * sizeof(int32_t)))
    return rc;
  int32_t* enc = static_cast<int32_t*>(p->hub_colind.p);
  const int sweep_grid =
      int(std::max<int64_t>(1, std::min<int64_t>((nnz + pad + 255) / 256, int64_t(p->num_sms) * 16)));

  std::vector<int32_t> hub; // the chosen columns, ascending
  int64_t refs = 0;
  Scratch t;
  if (cap > 0 && nnz > 0 && cols > 0) {
    if (int rc = reserve(p, t.counts, size_t(cols) * sizeof(int)))
      return rc;
    int* counts = static_cast<int*>(t.counts.p);
    B200_CUDA_TRY(p, cudaMemsetAsync(counts, 0, size_t(cols) * sizeof(int), s));
    hub_count_kernel<<<sweep_grid, 256, 0, s>>>(colind, nnz, cols, counts);
    if (int rc = launch_check(p, "hub_count_kernel"))
      return rc;
    // candidates: columns referenced at least min_count times, in ascending order
    // shared-memory table: a hub costs one load per CTA and launch; global table: one
    // gather per launch, repaid from the second reference on
    const int min_count = int(std::min<int64_t>(
        p->hub_min_count > 0 ? p->hub_min_count
                             : (by_popularity ? int64_t(3) : 2 * int64_t(p->num_sms)),
        0x7fffffff));
    if (int rc = reserve(p, t.cand, size_t(cols) * sizeof(int32_t)))
      return rc;
    if (int rc = reserve(p, t.nsel, sizeof(int64_t)))
      return rc;
    int32_t* cand = static_cast<int32_t*>(t.cand.p);
    int64_t* d_nsel = static_cast<int64_t*>(t.nsel.p);
    cub::CountingInputIterator<int32_t> ids(0);
    const AtLeast pred{counts, min_count};
    size_t ws_bytes = 0;
    B200_CUDA_TRY(p, cub::DeviceSelect::If(nullptr, ws_bytes, ids, cand, d_nsel, int(cols), pred, s));
    if (int rc = reserve(p, t.ws, ws_bytes))
      return rc;
    B200_CUDA_TRY(p, cub::DeviceSelect::If(t.ws.p, ws_bytes, ids, cand, d_nsel, int(cols), pred, s));
    int64_t nsel = 0;
    B200_CUDA_TRY(p, cudaMemcpyAsync(&nsel, d_nsel, sizeof(nsel), cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
    if (nsel > 0) {
      // (count descending, column ascending): the sort is stable and the candidates ascend
      if (int rc = reserve(p, t.cand_counts, size_t(nsel) * sizeof(int)))
        return rc;
      if (int rc = reserve(p, t.sorted_counts, size_t(nsel) * sizeof(int)))
        return rc;
      if (int rc = reserve(p, t.sorted_cols, size_t(nsel) * sizeof(int32_t)))
        return rc;
      int* cand_counts = static_cast<int*>(t.cand_counts.p);
      int* sorted_counts = static_cast<int*>(t.sorted_counts.p);
      int32_t* sorted_cols = static_cast<int32_t*>(t.sorted_cols.p);
      hub_gather_counts_kernel<<<unsigned((nsel + 255) / 256), 256, 0, s>>>(counts, cand, nsel,
                                                                           cand_counts);
      if (int rc = launch_check(p, "hub_gather_counts_kernel"))
        return rc;
      ws_bytes = 0;
      B200_CUDA_TRY(p, cub::DeviceRadixSort::SortPairsDescending(
                           nullptr, ws_bytes, cand_counts, sorted_counts, cand, sorted_cols,
                           int(nsel), 0, 32, s));
      if (int rc = reserve(p, t.ws, ws_bytes))
        return rc;
      B200_CUDA_TRY(p, cub::DeviceRadixSort::SortPairsDescending(
                           t.ws.p, ws_bytes, cand_counts, sorted_counts, cand, sorted_cols,
                           int(nsel), 0, 32, s));
      const int64_t h = std::min<int64_t>(nsel, cap);
      if (by_popularity) {
        // millions of hubs: everything stays on the device.  refs = sum of the top h counts
        cub::TransformInputIterator<int64_t, ToInt64, const int*> as64(sorted_counts, ToInt64());
        ws_bytes = 0;
        B200_CUDA_TRY(p, cub::DeviceReduce::Sum(nullptr, ws_bytes, as64, d_nsel, int(h), s));
        if (int rc = reserve(p, t.ws, ws_bytes))
          return rc;
        B200_CUDA_TRY(p, cub::DeviceReduce::Sum(t.ws.p, ws_bytes, as64, d_nsel, int(h), s));
        B200_CUDA_TRY(p, cudaMemcpyAsync(&refs, d_nsel, sizeof(refs), cudaMemcpyDeviceToHost, s));
        if (int rc = reserve(p, p->hub_cols, size_t(std::max<int64_t>(h, 1)) * sizeof(int32_t)))
          return rc;
        B200_CUDA_TRY(p, cudaMemcpyAsync(p->hub_cols.p, sorted_cols, size_t(h) * sizeof(int32_t),
                                         cudaMemcpyDeviceToDevice, s));
        B200_CUDA_TRY(p, cudaStreamSynchronize(s));
        B200_CUDA_TRY(p, cudaMemsetAsync(counts, 0xff, size_t(cols) * sizeof(int), s));
        if (h > 0) {
          hub_slot_kernel<<<unsigned((h + 255) / 256), 256, 0, s>>>(
              static_cast<const int32_t*>(p->hub_cols.p), h, counts);
          if (int rc = launch_check(p, "hub_slot_kernel"))
            return rc;
        }
        hub_encode_kernel<<<sweep_grid, 256, 0, s>>>(colind, nnz, cols, counts, pad, enc);
        if (int rc = launch_check(p, "hub_encode_kernel"))
          return rc;
        B200_CUDA_TRY(p, cudaStreamSynchronize(s));
        p->hub_count = h;
        p->hub_refs = refs;
        p->hub_state = 1;
        return SPBLAS_B200_SUCCESS;
      }
      std::vector<int> top_counts(size_t(h), 0);
      hub.resize(size_t(h));
      B200_CUDA_TRY(p, cudaMemcpyAsync(top_counts.data(), sorted_counts, size_t(h) * sizeof(int),
                                       cudaMemcpyDeviceToHost, s));
      B200_CUDA_TRY(p, cudaMemcpyAsync(hub.data(), sorted_cols, size_t(h) * sizeof(int32_t),
                                       cudaMemcpyDeviceToHost, s));
      B200_CUDA_TRY(p, cudaStreamSynchronize(s));
      refs = std::accumulate(top_counts.begin(), top_counts.end(), int64_t(0));
      // ascending columns: neighbouring hub slots are neighbouring elements of x, so the
      // CTA's load of the table reads whole lines where hubs cluster
      std::sort(hub.begin(), hub.end());
    }
    // slot_of[col] (reusing the counts): -1, or the hub's number
    const int64_t h = int64_t(hub.size());
    if (int rc = reserve(p, p->hub_cols, size_t(std::max<int64_t>(h, 1)) * sizeof(int32_t)))
      return rc;
    B200_CUDA_TRY(p, cudaMemsetAsync(counts, 0xff, size_t(cols) * sizeof(int), s));
    if (h > 0) {
      B200_CUDA_TRY(p, cudaMemcpyAsync(p->hub_cols.p, hub.data(), size_t(h) * sizeof(int32_t),
                                       cudaMemcpyHostToDevice, s));
      hub_slot_kernel<<<unsigned((h + 255) / 256), 256, 0, s>>>(
          static_cast<const int32_t*>(p->hub_cols.p), h, counts);
      if (int rc = launch_check(p, "hub_slot_kernel"))
        return rc;
    }
    hub_encode_kernel<<<sweep_grid, 256, 0, s>>>(colind, nnz, cols, counts, pad, enc);
    if (int rc = launch_check(p, "hub_encode_kernel"))
      return rc;
    // `hub` (pageable host memory) and the scratch buffers go away when this returns
    B200_CUDA_TRY(p, cudaStreamSynchronize(s));
  } else {
    if (int rc = reserve(p, p->hub_cols, sizeof(int32_t)))
      return rc;
    if (nnz + pad > 0) { // no analysis (capacity 0 or an empty matrix): a plain copy
      B200_CUDA_TRY(p, cudaMemsetAsync(enc, 0, size_t(pad) * sizeof(int32_t), s));
      B200_CUDA_TRY(p, cudaMemcpyAsync(enc + pad, colind, size_t(nnz) * sizeof(int32_t),
                                       cudaMemcpyDeviceToDevice, s));
    }
  }
  p->hub_count = int64_t(hub.size());
  p->hub_refs = refs;
  p->hub_state = 1;
  return SPBLAS_B200_SUCCESS;
}

} // namespace b200
