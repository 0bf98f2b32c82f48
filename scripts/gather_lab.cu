// gather_lab.cu — measurement aid (not part of the library): how fast can one SM gather
// 4- and 8-byte elements of x through an index stream, and does the path matter?
//
// DESIGN 4.13 reads the warp-stream kernel's plateau as "every gather in flight holds an L1
// line": ~0.74 gathers per cycle per SM on C4, the LDG probe at 0.92 of the request port.
// cp.async (LDGSTS) lands its data in shared memory; the .cg form does not allocate in L1 at
// all.  If the landing zone in shared memory replaces L1 as what bounds the misses in flight,
// a walk built on it has more of them.  Variants, all over the same colind (int32, 128-bit
// streaming loads) and values:
//   0  LDG      ld.global.nc, 8 gathers in flight per thread (the library's probe)
//   1  LDG.cg   the same through ld.global.cg
//   2  LDGSTS.ca 4/8 bytes per gather, G groups of 8 in flight per thread
//   3  LDGSTS.cg 16 bytes per gather (the aligned 16 bytes holding the element)
// Index distributions: uniform over n; "rmat": every bit of the index is 1 with probability
// 0.24 (R-MAT a,b,c,d = .57,.19,.19,.05 gives column bits b+d), which has C4's popularity skew.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/gather_lab scripts/gather_lab.cu
//   build/gather_lab            (prints one JSON line per measurement)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void fill_idx(int* idx, int64_t nnz, int bits, int rmat) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nnz; i += stride) {
    if (!rmat) {
      idx[i] = int(mix(uint64_t(i)) & ((1ull << bits) - 1));
    } else {
      uint32_t c = 0;
      uint64_t r = mix(uint64_t(i));
      for (int b = 0; b < bits; ++b) {
        if ((b & 7) == 0 && b)
          r = mix(r + uint64_t(b));
        c |= uint32_t(((r >> (8 * (b & 7))) & 255u) < 61u) << b; // 61/256 = 0.24
      }
      idx[i] = int(c);
    }
  }
}

template <typename T>
__global__ void fill_val(T* v, int64_t n, uint64_t seed) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    v[i] = T(double(mix(uint64_t(i) + seed) >> 11) * (1.0 / 9007199254740992.0));
}

__device__ __forceinline__ uint4 ld_stream_16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
template <typename T>
__device__ __forceinline__ void ld_vals(const T* p, T (&v)[4]) {
  if constexpr (sizeof(T) == 4) {
    uint4 r = ld_stream_16(p);
    v[0] = __uint_as_float(r.x), v[1] = __uint_as_float(r.y), v[2] = __uint_as_float(r.z),
    v[3] = __uint_as_float(r.w);
  } else {
    uint4 a = ld_stream_16(p), b = ld_stream_16(p + 2);
    v[0] = __longlong_as_double((long long)((uint64_t(a.y) << 32) | a.x));
    v[1] = __longlong_as_double((long long)((uint64_t(a.w) << 32) | a.z));
    v[2] = __longlong_as_double((long long)((uint64_t(b.y) << 32) | b.x));
    v[3] = __longlong_as_double((long long)((uint64_t(b.w) << 32) | b.z));
  }
}

// ---- variants 0/1: LDG -----------------------------------------------------------------------
// U quads of indices per thread and step: 4 U gathers in flight per thread
template <typename T, bool CG, int U>
__global__ void __launch_bounds__(256)
k_ldg(const int* __restrict__ idx, const T* __restrict__ val, const T* __restrict__ x,
      int64_t nnz, T* __restrict__ out) {
  const int64_t nq = nnz >> 2;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  T acc = T(0);
  int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; q + (U - 1) * stride < nq; q += U * stride) {
    int c[U][4];
    T v[U][4], xv[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint4 cc = ld_stream_16(idx + 4 * (q + u * stride));
      c[u][0] = int(cc.x), c[u][1] = int(cc.y), c[u][2] = int(cc.z), c[u][3] = int(cc.w);
      ld_vals(val + 4 * (q + u * stride), v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        xv[u][j] = CG ? __ldcg(x + c[u][j]) : __ldg(x + c[u][j]);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        acc += v[u][j] * xv[u][j];
  }
  out[int64_t(blockIdx.x) * blockDim.x + threadIdx.x] = acc;
}

// ---- variants 2/3: LDGSTS --------------------------------------------------------------------
// Every thread owns G groups x 8 landing slots of SLOT bytes.  Per step: the quads of indices
// for one group (8 gathers) are loaded, 8 cp.async issued and committed; the group issued G-1
// steps earlier is waited for, read and multiplied.  values are re-loaded at consumption
// (they are streamed, not gathered: their latency hides behind the G-1 groups in flight).
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES)
                 : "memory");
}
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, int SLOT, int G, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_ldgsts(const int* __restrict__ idx, const T* __restrict__ val, const T* __restrict__ x,
         int64_t nnz, T* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  // slot (g, j) of thread t: conflict-free for the 16-byte form (consecutive threads,
  // consecutive 16-byte slots), 4-byte form likewise
  const uint32_t base = uint32_t(__cvta_generic_to_shared(smem));
  auto slot = [&](int g, int j) -> uint32_t {
    return base + uint32_t(((g * 8 + j) * THREADS + threadIdx.x) * SLOT);
  };
  const int64_t nq = nnz >> 3; // units of 8 gathers
  const int64_t stride = int64_t(gridDim.x) * THREADS;
  int64_t q = int64_t(blockIdx.x) * THREADS + threadIdx.x;
  T acc = T(0);
  int cidx[G][8];            // (only the low bits are needed for the 16-byte form)
  int64_t qs[G];
  int issued = 0, done = 0;
  // number of groups this thread will process
  int64_t mine = q < nq ? (nq - q + stride - 1) / stride : 0;
  auto issue = [&](int g, int64_t qq) {
    const uint4 c0 = ld_stream_16(idx + 8 * qq), c1 = ld_stream_16(idx + 8 * qq + 4);
    const int c[8] = {int(c0.x), int(c0.y), int(c0.z), int(c0.w),
                      int(c1.x), int(c1.y), int(c1.z), int(c1.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cidx[g][j] = c[j];
      if constexpr (SLOT == 16) {
        const char* src = reinterpret_cast<const char*>(x + c[j]);
        src = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15));
        cp_async<16>(slot(g, j), src);
      } else {
        cp_async<sizeof(T)>(slot(g, j), x + c[j]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    qs[g] = qq;
  };
  auto consume = [&](int g) {
    T v0[4], v1[4];
    ld_vals(val + 8 * qs[g], v0);
    ld_vals(val + 8 * qs[g] + 4, v1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t a = slot(g, j);
      if constexpr (SLOT == 16)
        a += (uint32_t(cidx[g][j]) * uint32_t(sizeof(T))) & 15u;
      T xv;
      if constexpr (sizeof(T) == 4) {
        uint32_t r;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(a));
        xv = __uint_as_float(r);
      } else {
        unsigned long long r;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(a));
        xv = __longlong_as_double((long long)r);
      }
      acc += (j < 4 ? v0[j] : v1[j - 4]) * xv;
    }
  };
  // prologue: G-1 groups in flight
#pragma unroll
  for (int g = 0; g < G - 1; ++g) {
    if (issued < mine) {
      issue(g, q + int64_t(issued) * stride);
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    ++issued;
  }
  // steady state, unrolled by G so that the group index is a compile-time constant
  while (done < mine) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int gi = (g + G - 1) % G; // the slot freed by the previous consume
      if (issued < mine) {
        issue(gi, q + int64_t(issued) * stride);
      } else {
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      ++issued;
      cp_wait<G - 1>();
      if (done < mine)
        consume(g);
      ++done;
    }
  }
  cp_wait<0>();
  out[int64_t(blockIdx.x) * THREADS + threadIdx.x] = acc;
}

struct Result {
  double ms;
  double checksum;
};

template <typename F>
Result time_it(F launch, void* out, size_t out_elems, size_t elem, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i)
    launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i)
    launch();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned char> h(out_elems * elem);
  CK(cudaMemcpy(h.data(), out, h.size(), cudaMemcpyDeviceToHost));
  double cs = 0;
  for (size_t i = 0; i < out_elems; ++i)
    cs += elem == 4 ? double(reinterpret_cast<float*>(h.data())[i])
                    : reinterpret_cast<double*>(h.data())[i];
  return {ms / reps, cs};
}

template <typename T>
void report(const char* name, const char* dist, int bits, int64_t nnz, int ctas, int threads,
            int smem, Result r, int sms, double mhz) {
  const double cyc = r.ms * 1e-3 * mhz * 1e6;
  printf("{\"variant\": \"%s\", \"T\": %d, \"dist\": \"%s\", \"x_bits\": %d, \"nnz\": %lld, "
         "\"ctas_per_sm\": %d, \"threads\": %d, \"smem\": %d, \"ms\": %.4f, "
         "\"gathers_per_cycle_per_sm\": %.3f, \"checksum\": %.6e}\n",
         name, int(sizeof(T)), dist, bits, (long long)nnz, ctas, threads, smem, r.ms,
         double(nnz) / cyc / sms, r.checksum);
  fflush(stdout);
}

template <typename T, int SLOT, int G, int THREADS>
void run_ldgsts(const char* name, const char* dist, int bits, const int* idx, const T* val,
                const T* x, int64_t nnz, T* out, int sms, double mhz, int ctas) {
  const int smem = G * 8 * THREADS * SLOT;
  auto kern = k_ldgsts<T, SLOT, G, THREADS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  if (occ < 1)
    return;
  if (ctas > occ)
    ctas = occ;
  const int grid = sms * ctas;
  Result r = time_it([&] { kern<<<grid, THREADS, smem>>>(idx, val, x, nnz, out); }, out,
                     size_t(grid) * THREADS, sizeof(T), 10);
  report<T>(name, dist, bits, nnz, ctas, THREADS, smem, r, sms, mhz);
}

template <typename T>
void run_all(const char* dist, int bits, int rmat, int64_t nnz, int sms, double mhz) {
  int* idx;
  T *val, *x, *out;
  const int64_t n = int64_t(1) << bits;
  CK(cudaMalloc(&idx, nnz * 4));
  CK(cudaMalloc(&val, nnz * sizeof(T)));
  CK(cudaMalloc(&x, n * sizeof(T)));
  CK(cudaMalloc(&out, size_t(sms) * 16 * 1024 * sizeof(T)));
  fill_idx<<<sms * 8, 256>>>(idx, nnz, bits, rmat);
  fill_val<T><<<sms * 8, 256>>>(val, nnz, 1);
  fill_val<T><<<sms * 8, 256>>>(x, n, 2);
  CK(cudaDeviceSynchronize());
  for (int ctas : {4, 6, 8}) {
    const int grid = sms * ctas;
    Result r = time_it([&] { k_ldg<T, false, 2><<<grid, 256>>>(idx, val, x, nnz, out); }, out,
                       size_t(grid) * 256, sizeof(T), 10);
    report<T>("ldg x8", dist, bits, nnz, ctas, 256, 0, r, sms, mhz);
  }
  for (int ctas : {3, 4, 6}) {
    const int grid = sms * ctas;
    Result r = time_it([&] { k_ldg<T, false, 4><<<grid, 256>>>(idx, val, x, nnz, out); }, out,
                       size_t(grid) * 256, sizeof(T), 10);
    report<T>("ldg x16", dist, bits, nnz, ctas, 256, 0, r, sms, mhz);
  }
  {
    const int grid = sms * 8;
    Result r = time_it([&] { k_ldg<T, true, 2><<<grid, 256>>>(idx, val, x, nnz, out); }, out,
                       size_t(grid) * 256, sizeof(T), 10);
    report<T>("ldg.cg x8", dist, bits, nnz, 8, 256, 0, r, sms, mhz);
  }
  // LDGSTS.ca, element-sized landing slots
  run_ldgsts<T, sizeof(T), 2, 256>("ldgsts.ca G2", dist, bits, idx, val, x, nnz, out, sms, mhz, 8);
  run_ldgsts<T, sizeof(T), 4, 256>("ldgsts.ca G4", dist, bits, idx, val, x, nnz, out, sms, mhz, 8);
  run_ldgsts<T, sizeof(T), 4, 256>("ldgsts.ca G4", dist, bits, idx, val, x, nnz, out, sms, mhz, 4);
  run_ldgsts<T, sizeof(T), 8, 256>("ldgsts.ca G8", dist, bits, idx, val, x, nnz, out, sms, mhz, 4);
  // LDGSTS.cg, 16-byte landing slots (no L1 allocation)
  run_ldgsts<T, 16, 2, 256>("ldgsts.cg16 G2", dist, bits, idx, val, x, nnz, out, sms, mhz, 3);
  run_ldgsts<T, 16, 3, 256>("ldgsts.cg16 G3", dist, bits, idx, val, x, nnz, out, sms, mhz, 2);
  run_ldgsts<T, 16, 4, 256>("ldgsts.cg16 G4", dist, bits, idx, val, x, nnz, out, sms, mhz, 1);
  run_ldgsts<T, 16, 2, 512>("ldgsts.cg16 G2 t512", dist, bits, idx, val, x, nnz, out, sms, mhz, 1);
  run_ldgsts<T, 16, 3, 512>("ldgsts.cg16 G3 t512", dist, bits, idx, val, x, nnz, out, sms, mhz, 1);
  CK(cudaFree(idx));
  CK(cudaFree(val));
  CK(cudaFree(x));
  CK(cudaFree(out));
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const double mhz = khz / 1000.0;
  fprintf(stderr, "SMs %d, clock %.0f MHz (nominal max; gathers/cycle assume it)\n", sms, mhz);
  // C1-like: x = 4 MB (fp32), uniform; C4-like: x = 64 MB (fp32), skewed and uniform;
  // C5-like per GPU: fp64, x = 128 MB .. beyond L2
  run_all<float>("uniform", 20, 0, int64_t(1) << 25, sms, mhz);
  run_all<float>("rmat", 24, 1, int64_t(1) << 27, sms, mhz);
  run_all<float>("uniform", 24, 0, int64_t(1) << 27, sms, mhz);
  run_all<double>("rmat", 24, 1, int64_t(1) << 27, sms, mhz);
  return 0;
}
