// host_exec.cu — y_host = alpha * A * x_host with HOST operands (spblas_b200_spmv_host).
//
// The reference's GPU backends take device pointers only
// (include/spblas/vendor/cusparse/spmv_impl.hpp:50-66); a caller whose vectors live in
// host memory pays upload + product + download in sequence.  Here the three overlap:
// the partition's tiles are cut into chunks, and for every chunk
//     copy stream   x_host[frontier .. need_c)  ->  d_x      (only what chunk c adds)
//     plan stream   tiles of chunk c (SpMV kernel + carry fix-up for the rows it ends)
//     copy stream 2 d_y[rows of chunk c]        ->  y_host
// chained by events, so PCIe runs in both directions while the kernels work.  need_c is
// the running maximum of the columns referenced by chunks 0..c (found once per
// structure by a reduction over colind): a banded matrix uploads x progressively, a
// matrix with scattered columns degrades to "upload all, then multiply" with the
// download still overlapped.  The result is bit-identical to spblas_b200_spmv: the
// same kernels run on the same tiles, only the launch boundaries move.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

namespace {

constexpr int kMaxHostChunks = 64;

// max column index of each chunk's nonzeros (and the smallest column of all, in
// colmax[gridDim.x]): grid (chunks, slices)
template <typename I>
__global__ void __launch_bounds__(256)
chunk_colmax_kernel(const I* __restrict__ colind, const int64_t* __restrict__ tile_starts,
                    const int64_t* __restrict__ chunk_tile, long long* __restrict__ colmax) {
  const int c = blockIdx.x;
  const int64_t k0 = tile_starts[2 * chunk_tile[c] + 1];
  const int64_t k1 = tile_starts[2 * chunk_tile[c + 1] + 1];
  long long best = -1, least = INT64_MAX;
  for (int64_t k = k0 + int64_t(blockIdx.y) * blockDim.x + threadIdx.x; k < k1;
       k += int64_t(gridDim.y) * blockDim.x) {
    const long long v = static_cast<long long>(ld_stream(colind + k));
    best = v > best ? v : best;
    least = v < least ? v : least;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
    const long long other2 = __shfl_xor_sync(0xffffffffu, least, o);
    least = other2 < least ? other2 : least;
  }
  if ((threadIdx.x & 31) == 0 && best >= 0) {
    atomicMax(colmax + c, best);
    atomicMin(colmax + gridDim.x, least);
  }
}

int build_chunks(spblas_b200_plan* p, const int64_t* ts, int64_t units) {
  int chunks = p->host_chunks_override > 0 ? p->host_chunks_override : 16;
  chunks = std::min<int64_t>(std::min(chunks, kMaxHostChunks), std::max<int64_t>(units, 1));
  p->hc_tile.assign(chunks + 1, 0);
  p->hc_row.assign(chunks + 1, 0);
  p->hc_xneed.assign(chunks, 0);
  for (int c = 0; c <= chunks; ++c)
    p->hc_tile[c] = units * c / chunks;
  if (units == 0) {
    p->hc_row[chunks] = p->csr_rows;
    p->host_chunks = chunks;
    return SPBLAS_B200_SUCCESS;
  }
  // device scratch: chunk_tile table followed by the per-chunk maxima
  const size_t bytes = size_t(2 * chunks + 2) * sizeof(int64_t);
  if (int rc = reserve(p, p->hc_colmax, bytes))
    return rc;
  int64_t* d_tile = static_cast<int64_t*>(p->hc_colmax.p);
  long long* d_max = reinterpret_cast<long long*>(d_tile + chunks + 1);
  B200_CUDA_TRY(p, cudaMemcpyAsync(d_tile, p->hc_tile.data(), size_t(chunks + 1) * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, p->stream));
  B200_CUDA_TRY(p, cudaMemsetAsync(d_max, 0xff, size_t(chunks) * sizeof(long long), p->stream));
  const long long huge = INT64_MAX;
  B200_CUDA_TRY(p, cudaMemcpyAsync(d_max + chunks, &huge, sizeof(huge), cudaMemcpyHostToDevice,
                                   p->stream));
  const dim3 grid(chunks, 64);
  if (p->idx_type == SPBLAS_B200_I64)
    chunk_colmax_kernel<int64_t><<<grid, 256, 0, p->stream>>>(
        static_cast<const int64_t*>(p->csr_colind), ts, d_tile, d_max);
  else
    chunk_colmax_kernel<int32_t><<<grid, 256, 0, p->stream>>>(
        static_cast<const int32_t*>(p->csr_colind), ts, d_tile, d_max);
  B200_CUDA_TRY(p, cudaGetLastError());
  std::vector<long long> h_max(chunks + 1);
  B200_CUDA_TRY(p, cudaMemcpyAsync(h_max.data(), d_max, size_t(chunks + 1) * sizeof(long long),
                                   cudaMemcpyDeviceToHost, p->stream));
  // first row of every chunk: the row coordinate of its first tile
  std::vector<int64_t> h_row(chunks + 1);
  for (int c = 0; c <= chunks; ++c)
    B200_CUDA_TRY(p, cudaMemcpyAsync(&h_row[c], ts + 2 * p->hc_tile[c], sizeof(int64_t),
                                     cudaMemcpyDeviceToHost, p->stream));
  B200_CUDA_TRY(p, cudaStreamSynchronize(p->stream));
  int64_t need = 0;
  for (int c = 0; c < chunks; ++c) {
    need = std::max<int64_t>(need, h_max[c] + 1);
    p->hc_xneed[c] = std::min<int64_t>(need, p->csr_cols);
    p->hc_row[c] = h_row[c];
  }
  p->hc_row[chunks] = p->csr_rows;
  // columns below the smallest referenced one are never uploaded (a row-block shard of
  // a banded matrix touches only its own band of x)
  p->hc_xlo = std::min<int64_t>(h_max[chunks] == INT64_MAX ? 0 : h_max[chunks], need);
  if (!p->h2d_stream)
    B200_CUDA_TRY(p, cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
  if (!p->d2h_stream)
    B200_CUDA_TRY(p, cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
  while (p->hc_events.size() < size_t(2 * chunks + 2)) {
    cudaEvent_t ev;
    B200_CUDA_TRY(p, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    p->hc_events.push_back(ev);
  }
  p->host_chunks = chunks;
  return SPBLAS_B200_SUCCESS;
}

} // namespace

void release_host_exec(spblas_b200_plan* p) {
  for (cudaEvent_t ev : p->hc_events)
    cudaEventDestroy(ev);
  p->hc_events.clear();
  if (p->h2d_stream)
    cudaStreamDestroy(p->h2d_stream);
  if (p->d2h_stream)
    cudaStreamDestroy(p->d2h_stream);
  p->h2d_stream = p->d2h_stream = nullptr;
  p->host_chunks = 0;
}

int run_spmv_host(spblas_b200_plan* p, int val_type, const void* alpha,
                  const void* values, const void* h_x, void* h_y, void* d_x,
                  void* d_y) {
  if (p->scatter.n > 0 || p->barrier.n > 0)
    return fail(p, SPBLAS_B200_NOT_SUPPORTED,
                "host-buffer execute on a plan with a fused exchange");
  // the chunks are cut at units of the partition the chosen kernel runs on (tiles, or
  // warp streams), so every unit is processed exactly as in a device-vector execute
  int variant = 0;
  const int64_t* starts = nullptr;
  int64_t units = 0;
  // (the hub variant reloads its shared-memory table of x on every launch, and a chunk's
  // launch would read parts of x that have not arrived yet: chunks take the plain walk)
  struct Active {
    spblas_b200_plan* p;
    explicit Active(spblas_b200_plan* q) : p(q) { p->host_exec_active = true; }
    ~Active() { p->host_exec_active = false; }
  } active(p);
  if (int rc = prepare_spmv(p, val_type, values, &variant, &starts, &units))
    return rc;
  if (p->host_chunks == 0 || p->hc_variant != variant) {
    if (int rc = build_chunks(p, starts, units))
      return rc;
    p->hc_variant = variant;
  }
  const int chunks = p->host_chunks;
  const size_t sT = type_size_val(val_type);
  p->last_launches = 0;
  if (units == 0) {
    // no rows end anywhere: y (if any) is all zeros
    if (p->csr_rows > 0)
      std::memset(h_y, 0, size_t(p->csr_rows) * sT);
    return SPBLAS_B200_SUCCESS;
  }
  cudaEvent_t ev_start = p->hc_events[2 * chunks], ev_done = p->hc_events[2 * chunks + 1];
  // the copies may not overtake work already queued on the plan's stream that uses
  // d_x / d_y (an earlier execute)
  B200_CUDA_TRY(p, cudaEventRecord(ev_start, p->stream));
  B200_CUDA_TRY(p, cudaStreamWaitEvent(p->h2d_stream, ev_start, 0));
  B200_CUDA_TRY(p, cudaStreamWaitEvent(p->d2h_stream, ev_start, 0));
  int64_t frontier = p->hc_xlo;
  for (int c = 0; c < chunks; ++c) {
    cudaEvent_t ev_up = p->hc_events[2 * c], ev_mul = p->hc_events[2 * c + 1];
    if (p->hc_xneed[c] > frontier) {
      B200_CUDA_TRY(p, cudaMemcpyAsync(static_cast<char*>(d_x) + size_t(frontier) * sT,
                                       static_cast<const char*>(h_x) + size_t(frontier) * sT,
                                       size_t(p->hc_xneed[c] - frontier) * sT,
                                       cudaMemcpyHostToDevice, p->h2d_stream));
      frontier = p->hc_xneed[c];
      B200_CUDA_TRY(p, cudaEventRecord(ev_up, p->h2d_stream));
      B200_CUDA_TRY(p, cudaStreamWaitEvent(p->stream, ev_up, 0));
    }
    if (int rc = run_spmv_tiles(p, val_type, alpha, values, d_x, d_y, p->hc_tile[c],
                                p->hc_tile[c + 1]))
      return rc;
    const int64_t r0 = p->hc_row[c], r1 = p->hc_row[c + 1];
    if (r1 > r0) {
      B200_CUDA_TRY(p, cudaEventRecord(ev_mul, p->stream));
      B200_CUDA_TRY(p, cudaStreamWaitEvent(p->d2h_stream, ev_mul, 0));
      B200_CUDA_TRY(p, cudaMemcpyAsync(static_cast<char*>(h_y) + size_t(r0) * sT,
                                       static_cast<const char*>(d_y) + size_t(r0) * sT,
                                       size_t(r1 - r0) * sT, cudaMemcpyDeviceToHost,
                                       p->d2h_stream));
    }
  }
  // the call is complete, in stream order, when the last download is
  B200_CUDA_TRY(p, cudaEventRecord(ev_done, p->d2h_stream));
  B200_CUDA_TRY(p, cudaStreamWaitEvent(p->stream, ev_done, 0));
  return SPBLAS_B200_SUCCESS;
}

} // namespace b200
