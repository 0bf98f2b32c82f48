// probe.cu — measurement aid, not part of the multiply path: the fastest this GPU can
// stream an index array and gather x through it.
//
// SpMV on a matrix with scattered columns is not bound by HBM bytes but by the gathers
// of x: every gathered element is its own 128-byte line for the L1 tag stage and its own
// 32-byte sector for L2.  The compulsory-bytes roofline cannot see that
// limit, so bench.py also reports each such workload against THIS kernel run on the
// workload's own colind (and values): the same loads as SpMV — 128-bit streaming loads
// of colind and values, one read-only gather of x per nonzero, eight in flight per
// thread — and nothing else: no rows, no shared memory, no reduction beyond one
// accumulator per thread.  Whatever SpMV kernel is written, it cannot beat this.
#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {
namespace {

template <typename T, typename I>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? 4 : 6)
gather_probe_kernel(const I* __restrict__ colind, const T* __restrict__ values,
                    const T* __restrict__ x, int64_t nnz, T* __restrict__ out) {
  const int64_t nq = nnz >> 2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  T acc = T(0);
  int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; q + stride < nq; q += 2 * stride) {
    const Quad<I> c0 = ld_stream_quad(colind + 4 * q);
    const Quad<I> c1 = ld_stream_quad(colind + 4 * (q + stride));
    Quad<T> v0, v1;
    if (values) {
      v0 = ld_stream_quad(values + 4 * q);
      v1 = ld_stream_quad(values + 4 * (q + stride));
    }
    T x0[4], x1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      x0[j] = ld_ro(x + c0.v[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      x1[j] = ld_ro(x + c1.v[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc += values ? v0.v[j] * x0[j] + v1.v[j] * x1[j] : x0[j] + x1[j];
  }
  if (q < nq) {
    const Quad<I> c0 = ld_stream_quad(colind + 4 * q);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      acc += ld_ro(x + c0.v[j]) * (values ? values[4 * q + j] : T(1));
  }
  out[int64_t(blockIdx.x) * blockDim.x + threadIdx.x] = acc;
}

template <typename T>
int probe_typed(cudaStream_t s, int idx_type, int64_t nnz, const void* colind,
                const void* values, const void* x, void* out, int grid) {
  if (idx_type == SPBLAS_B200_I64)
    gather_probe_kernel<T, int64_t><<<grid, 256, 0, s>>>(
        static_cast<const int64_t*>(colind), static_cast<const T*>(values),
        static_cast<const T*>(x), nnz, static_cast<T*>(out));
  else
    gather_probe_kernel<T, int32_t><<<grid, 256, 0, s>>>(
        static_cast<const int32_t*>(colind), static_cast<const T*>(values),
        static_cast<const T*>(x), nnz, static_cast<T*>(out));
  return cudaGetLastError() == cudaSuccess ? SPBLAS_B200_SUCCESS : SPBLAS_B200_CUDA_ERROR;
}

} // namespace
} // namespace b200

extern "C" int spblas_b200_probe_gather(void* cuda_stream, int idx_type, int val_type,
                                        int64_t nnz, const void* d_colind,
                                        const void* d_values, const void* d_x,
                                        void* d_out, int ctas_per_sm) {
  using namespace b200;
  if (!d_colind || !d_x || !d_out || nnz < 0)
    return SPBLAS_B200_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(d_colind) & 15u) ||
      (d_values && (reinterpret_cast<uintptr_t>(d_values) & 15u)))
    return SPBLAS_B200_INVALID_ARGUMENT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (ctas_per_sm < 1 || ctas_per_sm > 8)
    ctas_per_sm = 8;
  const int grid = sms * ctas_per_sm;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  switch (val_type) {
  case SPBLAS_B200_F32:
    return probe_typed<float>(s, idx_type, nnz, d_colind, d_values, d_x, d_out, grid);
  case SPBLAS_B200_F64:
    return probe_typed<double>(s, idx_type, nnz, d_colind, d_values, d_x, d_out, grid);
  default:
    return SPBLAS_B200_NOT_SUPPORTED;
  }
}
