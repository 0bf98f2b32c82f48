// spmv_kernels.cuh — the SpMV kernels and their launcher, as templates.
//
// Included by exactly two translation units: spmv.cu instantiates the 3-argument product
// (ADD = false), spmv_axpby.cu the 4-argument one (ADD = true: y = alpha A x + beta d), so that
// the two sets compile in parallel and the plain product carries no trace of the addend.
// (originally spmv.cu) — y = alpha * A * x for the effective CSR structure of a plan.
//
// Replaces the serial double loop of the reference
// (include/spblas/algorithms/multiply_impl.hpp:43-52: zero y, then
//  y[i] += a_ik * x[k] in storage order) and the cusparseSpMV call of the
// reference's NVIDIA backend (include/spblas/vendor/cusparse/spmv_impl.hpp:80-84).
//
// Work decomposition: merge-path tiles.  The inspect phase cut the merged
// sequence (row ends ++ nonzeros) into tiles of tile_items items, so every tile
// streams the same number of bytes whatever the row-length distribution (uniform
// 5-point stencil rows, Poisson(10) rows, R-MAT hubs).  A tile's nonzero range
// [k0, k1) starts anywhere; the kernels fetch the 16-byte aligned superset
// [k0 & ~3, ceil4(k1)) (at most 6 extra elements per tile) and index shared memory
// from the aligned origin, so all bulk traffic is 128-bit aligned.
//
// Two kernels consume that partition:
//
//  spmv_pipe_kernel (default) — persistent CTAs, each owning a contiguous run of
//    tiles.  One producer warp runs ahead of eight consumer warps through a ring of
//    shared-memory stages: per tile it issues two TMA bulk copies
//    (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) for the
//    colind and values ranges and element-wise cp.async for the row-end offsets
//    (any alignment), all tracked by the stage's "full" mbarrier.  Consumers wait
//    on the barrier, reduce complete rows straight out of the stage (thread-per-row
//    for short rows, in storage order like the reference; warp-per-row with a
//    shuffle reduction for long rows), gather x through the read-only path, write y
//    once, and release the stage through its "empty" mbarrier.  HBM latency is
//    hidden by the (stages-1) tiles in flight per CTA instead of by occupancy.
//
//  spmv_merge_tile_kernel (fallback; also selectable for A/B runs) — one tile per
//    CTA, register-staged 128-bit loads, products through shared memory.  Handles
//    colind/values that are not 16-byte aligned.
//
// In both, the tile's trailing partial row becomes a carry (row, value); a second
// tiny kernel adds carries to y in tile order (deterministic, no floating-point
// atomics).  y is written exactly once per row by the tile that holds the row's
// end, so beta = 0 semantics (stale y, even NaN, is discarded) hold without a memset.
#include <algorithm>

#pragma once

#include "device_utils.cuh"
#include "plan.hpp"

namespace b200 {

// 128-bit loads and bulk copies need 16-byte aligned colind / values (/ permutation)
inline int spmv_vec_ok(const spblas_b200_plan* p, const void* values) {
  const auto aligned16 = [](const void* q) {
    return (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  };
  const bool perm = p->csr_perm != nullptr;
  return aligned16(p->csr_colind) && (perm || aligned16(values)) &&
         (!perm || aligned16(p->csr_perm));
}

namespace {

template <typename T>
struct alignas(16) Vec4 {
  T v[4];
};

constexpr int kLongRow = 64; // thread-per-row tiles hand rows longer than this to warps
constexpr int kTileSlack = 16; // a tile can exceed tile_items by < 8 nonzeros

// ============================================================================
// mbarrier / TMA bulk-copy / cp.async primitives (PTX; sm_90+ features used on
// sm_100a).  Shared addresses are 32-bit shared-window addresses.
// ============================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
// src, dst and bytes are multiples of 16.  The data is streamed once: L2 evict_first.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src,
                                            uint32_t bytes, uint32_t bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int BYTES>
__device__ __forceinline__ void cp_async_small(uint32_t dst, const void* src) {
  static_assert(BYTES == 4 || BYTES == 8, "element-wise cp.async of 4 or 8 bytes");
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src),
               "n"(BYTES)
               : "memory");
}

// arrive on `bar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar)
               : "memory");
}

__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ============================================================================
// Fused exchange: the kernels that write y also write the rows a peer GPU needs
// straight into that peer's replica of the next x (NVLink peer stores, or one
// multimem store that the switch delivers to every GPU), and the carry fix-up
// kernel ends with the cross-GPU flag barrier — an iteration y -> x is the same two
// launches as on one GPU, with no collective call.
// ============================================================================
template <typename T>
struct ScatterArgs {
  int n;
  int multicast;
  int64_t lo_min, hi_max; // union of the ranges: tiles outside skip all checks
  T* dst[kMaxPeers];
  int64_t lo[kMaxPeers], hi[kMaxPeers];
  // y = alpha A x + beta d (the 4-argument multiply, spblas_b200_spmv_axpby): the addend is
  // fused into the one store every row gets; d == nullptr: beta = 0, nothing is read.  d may
  // alias y (each element is read, then written, by the same thread).  These live in the
  // kernel's parameter bank: the test is uniform and costs no register in the hot loops.
  const T* d;
  T beta;
  // a call on a cached one-shot structure: run only if the verify kernel that precedes it on
  // the stream found the structure unchanged (*gate == gate_value); nullptr: no gate
  const unsigned int* gate;
  unsigned int gate_value;
  // late push (fix-up kernel only): the rows are copied to the peers by the fix-up kernel's
  // last block, after every carry is in, instead of being stored by the product kernel
  int late;
};

template <typename T>
__device__ __forceinline__ bool gate_closed(const ScatterArgs<T>& sc) {
  return sc.gate != nullptr && *reinterpret_cast<const volatile unsigned int*>(sc.gate) != sc.gate_value;
}

// ADD is a template parameter of every kernel: the 3-argument product must not pay for the
// addend (measured: as a run-time test of sc.d it cost the pipelined kernel 16 % on C2,
// 0.228 -> 0.266 ms — profiles/r02_c2_epilogue_ab.txt); the ADD = true kernels are compiled in
// their own translation unit (spmv_axpby.cu).
template <bool ADD, typename T>
__device__ __forceinline__ T with_addend(const ScatterArgs<T>& sc, int64_t row, T v) {
  if constexpr (ADD)
    return v + sc.beta * sc.d[row];
  else
    return v;
}

struct BarrierArgs {
  int n;
  unsigned long long epoch;
  unsigned long long timeout_ns; // how long a rank waits for a peer before it gives up
  unsigned long long* remote[kMaxPeers];
  const unsigned long long* local[kMaxPeers];
  unsigned int* state;   // [0] blocks done
  unsigned int* gave_up; // host-mapped: set when a wait timed out (the next execute fails)
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <typename T>
__device__ __forceinline__ void multimem_store(T* p, T v) {
  if constexpr (sizeof(T) == 8) {
    asm volatile("multimem.st.global.f64 [%0], %1;" ::"l"(p),
                 "d"(*reinterpret_cast<const double*>(&v))
                 : "memory");
  } else {
    asm volatile("multimem.st.global.f32 [%0], %1;" ::"l"(p),
                 "f"(*reinterpret_cast<const float*>(&v))
                 : "memory");
  }
}

// does any destination want a row of [r_begin, r_end)?  Quick reject on the union of the
// ranges, then the ranges themselves: two peers that want the first and the last grid
// line of a block (halo of a banded matrix) have a union that spans the whole block, and
// only the tiles at its two ends may pay for the per-row checks.
template <typename T>
__device__ __forceinline__ bool wants_rows(const ScatterArgs<T>& sc, int64_t r_begin,
                                           int64_t r_end) {
  if (sc.n == 0 || r_begin >= sc.hi_max || r_end <= sc.lo_min)
    return false;
#pragma unroll 1
  for (int d = 0; d < sc.n; ++d)
    if (r_begin < sc.hi[d] && r_end > sc.lo[d])
      return true;
  return false;
}

template <typename T>
__device__ __forceinline__ void scatter_store(const ScatterArgs<T>& sc, int64_t row, T v) {
#pragma unroll 1
  for (int d = 0; d < sc.n; ++d) {
    if (row >= sc.lo[d] && row < sc.hi[d]) {
      if (sc.multicast)
        multimem_store(sc.dst[d] + row, v);
      else
        sc.dst[d][row] = v;
    }
  }
}

// ============================================================================
// Pipelined kernel
// ============================================================================
constexpr int kPipeHeaderBytes = 64;
constexpr int kPipeMaxStages = 8;
constexpr int kMaxUniformLen = 8; // longest row length with an unrolled exact path

struct PipeHeader {
  long long row0; // first row whose end lies in this tile
  long long kq0;  // absolute nonzero index of local slot 0 (k0 rounded down to 4)
  int nr;         // row ends in this tile
  int lo;         // local index of the tile's first nonzero (k0 - kq0)
  int hi;         // local index one past the tile's last nonzero (k1 - kq0)
  int nslots;     // local slots held by the stage (multiple of 4)
  int off_b;      // byte offset of the second staged array (values, or perm)
  int off_prod;   // byte offset of the products (== off_b unless a CSC image)
  int off_rowend; // byte offset of the row ends
  int uniform;    // L if rows row0+1 .. row0+nr-1 all have L entries (1..8), else 0
};
static_assert(sizeof(PipeHeader) <= kPipeHeaderBytes, "header too large");

// bytes one stage's data area must hold for a tile of `tile_items`: per nonzero
// colind + values (products overwrite the values in place); a CSC image stages
// colind + permutation and needs separate room for the products; a row end costs one
// offset.  Whichever is larger per merge item bounds the tile.
__host__ __device__ inline int pipe_stage_data_bytes(int tile_items, int sT, int sI,
                                                     int sO, bool perm) {
  const int per_nz = perm ? sI + sO + sT : sI + sT;
  const int per_item = per_nz > sO ? per_nz : sO;
  const int bytes = per_item * (tile_items + kTileSlack) + 128;
  return (bytes + 127) & ~127;
}

// ---- path 1: uniform tiles (stencils, fixed-degree graphs) ----------------------
// The inspect phase found that every complete row of the tile after the first has
// exactly L entries, so row r of those starts at e0 + L*r: no row-end lookups, no
// per-row branches.  One thread per row; all L column indices are read, all L
// gathers of x issued, then the FMAs run in storage order (the reference's order).
// The 32 lanes of a warp own 32 consecutive rows, so gather j of every lane falls in
// the same few 128-byte lines of x for a banded matrix: the gathers are coalesced.
template <int L, typename T, typename I>
__device__ __forceinline__ T dot_exact(const I* __restrict__ col,
                                       const T* __restrict__ val,
                                       const T* __restrict__ x, int b) {
  I c[L];
  T xv[L], av[L];
#pragma unroll
  for (int j = 0; j < L; ++j)
    c[j] = col[b + j];
#pragma unroll
  for (int j = 0; j < L; ++j)
    xv[j] = ld_ro(x + c[j]);
#pragma unroll
  for (int j = 0; j < L; ++j)
    av[j] = val[b + j];
  T sum = av[0] * xv[0];
#pragma unroll
  for (int j = 1; j < L; ++j)
    sum += av[j] * xv[j];
  return sum;
}

template <int L, int CONS, bool SCAT, bool ADD, typename T, typename I>
__device__ __forceinline__ void
uniform_rows(const I* __restrict__ col, const T* __restrict__ val,
             const T* __restrict__ x, T* __restrict__ yrow, const T alpha, int e0,
             int nrows, int tid, const ScatterArgs<T>& sc, int64_t rowbase) {
  for (int r = tid; r < nrows; r += CONS) {
    const T v = with_addend<ADD>(sc, rowbase + r, alpha * dot_exact<L, T, I>(col, val, x, e0 + L * r));
    yrow[r] = v;
    if constexpr (SCAT) // a tile a peer needs rows of (rare: the loop above stays lean)
      scatter_store(sc, rowbase + r, v);
  }
}

template <int CONS, bool SCAT, bool ADD, typename T, typename I>
__device__ __forceinline__ void
uniform_tile(int uni, const I* __restrict__ col, const T* __restrict__ val,
             const T* __restrict__ x, T* __restrict__ yrow, const T alpha, int e0,
             int nrows, int tid, const ScatterArgs<T>& sc, int64_t rowbase) {
  switch (uni) {
  case 1: uniform_rows<1, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  case 2: uniform_rows<2, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  case 3: uniform_rows<3, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  case 4: uniform_rows<4, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  case 5: uniform_rows<5, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  case 6: uniform_rows<6, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  case 7: uniform_rows<7, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  default: uniform_rows<8, CONS, SCAT, ADD, T, I>(col, val, x, yrow, alpha, e0, nrows, tid, sc, rowbase); break;
  }
}

// a row (or row fragment) [b, e) by one warp, straight from the staged operands
template <typename T, typename I>
__device__ __forceinline__ T dot_warp(const I* __restrict__ col,
                                      const T* __restrict__ val,
                                      const T* __restrict__ x, int b, int e, int lane) {
  T sum = T(0);
  int i = b + lane;
  for (; i + 32 < e; i += 64) {
    const I c0 = col[i], c1 = col[i + 32];
    const T x0 = ld_ro(x + c0), x1 = ld_ro(x + c1);
    sum += val[i] * x0;
    sum += val[i + 32] * x1;
  }
  if (i < e)
    sum += val[i] * ld_ro(x + col[i]);
  return warp_reduce_sum(sum);
}

// ---- path 2: general tiles -------------------------------------------------------
// (A) flat product phase: every consumer thread turns whole QUADS of staged entries
// into products a_ik * x_k with no knowledge of rows: two quads per thread are
// fetched from shared memory with 128-bit loads, their eight gathers of x are issued
// together through the read-only path, and the products are written back with
// 128-bit stores.  There are no per-element predicates: every slot of the stage holds
// a real matrix entry (entries of the neighbouring tiles in the aligned fringe; the
// producer zero-fills slots past the end of the arrays), and products outside
// [lo, hi) are simply never read.  (B) after a consumer-only barrier, rows are summed
// out of shared memory.
template <typename T, typename I>
__device__ __forceinline__ void quad_products(const I* __restrict__ col, const T* val,
                                              T* prod, const T* __restrict__ x, int qa,
                                              int qb, bool two) {
  const Vec4<I> ca = *reinterpret_cast<const Vec4<I>*>(col + 4 * qa);
  const Vec4<T> va = *reinterpret_cast<const Vec4<T>*>(val + 4 * qa);
  Vec4<I> cb = ca;
  Vec4<T> vb = va;
  if (two) {
    cb = *reinterpret_cast<const Vec4<I>*>(col + 4 * qb);
    vb = *reinterpret_cast<const Vec4<T>*>(val + 4 * qb);
  }
  T xa[4], xb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    xa[j] = ld_ro(x + ca.v[j]);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    xb[j] = ld_ro(x + cb.v[j]);
  Vec4<T> pa, pb;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pa.v[j] = va.v[j] * xa[j];
    pb.v[j] = vb.v[j] * xb[j];
  }
  *reinterpret_cast<Vec4<T>*>(prod + 4 * qa) = pa;
  if (two)
    *reinterpret_cast<Vec4<T>*>(prod + 4 * qb) = pb;
}

// CSC image: the stage holds the value permutation, the value itself is gathered
template <typename T, typename I, typename O>
__device__ __forceinline__ void quad_products_perm(const I* __restrict__ col,
                                                   const O* __restrict__ perm,
                                                   T* __restrict__ prod,
                                                   const T* __restrict__ values,
                                                   const T* __restrict__ x, int q) {
  const Vec4<I> c = *reinterpret_cast<const Vec4<I>*>(col + 4 * q);
  const Vec4<O> pi = *reinterpret_cast<const Vec4<O>*>(perm + 4 * q);
  T a[4], xv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] = ld_ro(values + pi.v[j]);
    xv[j] = ld_ro(x + c.v[j]);
  }
  Vec4<T> p;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    p.v[j] = a[j] * xv[j];
  *reinterpret_cast<Vec4<T>*>(prod + 4 * q) = p;
}

template <typename T>
__device__ __forceinline__ T prod_sum_thread(const T* __restrict__ prod, int b, int e) {
  T sum = T(0);
#pragma unroll 1
  for (int i = b; i < e; ++i) // storage order, like the reference's row loop
    sum += prod[i];
  return sum;
}

template <typename T>
__device__ __forceinline__ T prod_sum_warp(const T* __restrict__ prod, int b, int e,
                                           int lane) {
  T sum = T(0);
  for (int i = b + lane; i < e; i += 32)
    sum += prod[i];
  return warp_reduce_sum(sum);
}

template <typename T, typename I, typename O, int CW, bool ADD>
__global__ void __launch_bounds__(CW * 32 + 32, CW <= 8 ? 3 : 1)
spmv_pipe_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                 const T* __restrict__ values, const O* __restrict__ perm,
                 const T* __restrict__ x, T* __restrict__ y, const T alpha,
                 const int64_t* __restrict__ tile_starts,
                 const int* __restrict__ tile_uniform, const int64_t tile_first,
                 const int64_t num_tiles, const int64_t rows, const int64_t nnz_end,
                 int64_t* __restrict__ carry_row, T* __restrict__ carry_val,
                 const int stages, const int stage_data_bytes,
                 const __grid_constant__ ScatterArgs<T> sc) {
  constexpr int kPipeConsumerWarps = CW;        // consumer warps; warp CW is the producer
  constexpr int kPipeConsumers = CW * 32;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [full barriers][empty barriers][pad to 128][stage 0 header+data]...
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const int stage_bytes = kPipeHeaderBytes + stage_data_bytes;
  unsigned char* stage_base = smem + 128;
  __shared__ T s_red[kPipeConsumerWarps];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const bool has_perm = perm != nullptr;
  if (gate_closed(sc))
    return;

  // contiguous run of tiles for this CTA out of [tile_first, tile_first + num_tiles)
  const int64_t t_begin = tile_first + num_tiles * int64_t(blockIdx.x) / gridDim.x;
  const int64_t t_end = tile_first + num_tiles * int64_t(blockIdx.x + 1) / gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&bars[s]), 32 + 1);                 // full: 32 cp.async arrivals + expect_tx
      mbar_init(smem_u32(&bars[kPipeMaxStages + s]), kPipeConsumerWarps); // empty
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kPipeConsumerWarps) {
    // ======================= producer warp ======================================
    const uint64_t policy = policy_evict_first();
    int s = 0;
    uint32_t phase = 0;
    // tile coordinates are prefetched one tile ahead so that the table's load
    // latency never sits between two tiles' copies
    int64_t cur_row = 0, cur_k = 0, nxt_row = 0, nxt_k = 0;
    int cur_uni = 0, nxt_uni = 0;
    if (t_begin < t_end) {
      cur_row = tile_starts[2 * t_begin];
      cur_k = tile_starts[2 * t_begin + 1];
      nxt_row = tile_starts[2 * t_begin + 2];
      nxt_k = tile_starts[2 * t_begin + 3];
      cur_uni = has_perm ? 0 : tile_uniform[t_begin];
    }
    for (int64_t t = t_begin; t < t_end; ++t) {
      const int64_t row0 = cur_row, k0 = cur_k, row1 = nxt_row, k1 = nxt_k;
      const int uni = cur_uni;
      cur_row = nxt_row;
      cur_k = nxt_k;
      if (t + 1 < t_end) {
        nxt_row = tile_starts[2 * t + 4];
        nxt_k = tile_starts[2 * t + 5];
        nxt_uni = has_perm ? 0 : tile_uniform[t + 1];
      }
      cur_uni = nxt_uni;
      const int nr = int(row1 - row0);
      const int64_t kq0 = k0 & ~int64_t(3);             // aligned origin of the stage
      const int nslots = int(((k1 - kq0) + 3) & ~int64_t(3)); // local slots (multiple of 4)
      const int b_elem = has_perm ? int(sizeof(O)) : int(sizeof(T));
      const int off_b = nslots * int(sizeof(I));
      const int off_prod = has_perm ? off_b + nslots * int(sizeof(O)) : off_b;
      const int off_rowend = (off_prod + nslots * int(sizeof(T)) + 15) & ~15;
      // quads that lie entirely inside the arrays are bulk-copied; the arrays' last
      // partial quad (if any) is copied element-wise below
      const int64_t kfull = nnz_end & ~int64_t(3);
      int64_t kb1 = kq0 + nslots;
      if (kb1 > kfull)
        kb1 = kfull;
      const int bulk = kb1 > kq0 ? int(kb1 - kq0) : 0;

      unsigned char* st = stage_base + size_t(s) * stage_bytes;
      unsigned char* data = st + kPipeHeaderBytes;
      const uint32_t full = smem_u32(&bars[s]);
      const uint32_t empty = smem_u32(&bars[kPipeMaxStages + s]);

      if (lane == 0)
        mbar_wait(empty, phase ^ 1u); // stage free (passes at once on the first lap)
      __syncwarp();

      // Slots past the end of the arrays (last tile only) get column 0 and a zero
      // value / permutation 0, so that the predicate-free product phase gathers valid
      // addresses.  Plain stores: they must precede lane 0's releasing arrive below.
      const bool ragged_end = kq0 + nslots > kfull;
      if (ragged_end) {
        for (int li = bulk + lane; li < nslots; li += 32) {
          if (kq0 + li >= nnz_end) {
            reinterpret_cast<I*>(data)[li] = I(0);
            if (has_perm)
              reinterpret_cast<O*>(data + off_b)[li] = O(0);
            else
              reinterpret_cast<T*>(data + off_b)[li] = T(0);
          }
        }
        __syncwarp();
      }

      if (lane == 0) {
        PipeHeader* h = reinterpret_cast<PipeHeader*>(st);
        h->row0 = row0;
        h->kq0 = kq0;
        h->nr = nr;
        h->lo = int(k0 - kq0);
        h->hi = int(k1 - kq0);
        h->nslots = nslots;
        h->off_b = off_b;
        h->off_prod = off_prod;
        h->off_rowend = off_rowend;
        h->uniform = uni;
        mbar_arrive_expect_tx(full, uint32_t(bulk) * uint32_t(int(sizeof(I)) + b_elem));
        if (bulk > 0) {
          tma_load_1d(smem_u32(data), colind + kq0, uint32_t(bulk) * sizeof(I), full,
                      policy);
          if (!has_perm)
            tma_load_1d(smem_u32(data + off_b), values + kq0, uint32_t(bulk) * sizeof(T),
                        full, policy);
          else
            tma_load_1d(smem_u32(data + off_b), perm + kq0, uint32_t(bulk) * sizeof(O),
                        full, policy);
        }
      }
      // row ends: element-wise async copies (no alignment requirement).  A uniform
      // tile only needs the first one.
      {
        const uint32_t dst0 = smem_u32(data + off_rowend);
        const O* src0 = rowptr + row0 + 1;
        const int ncopy = uni > 0 ? (nr > 0 ? 1 : 0) : nr;
        for (int q = lane; q < ncopy; q += 32)
          cp_async_small<sizeof(O)>(dst0 + q * uint32_t(sizeof(O)), src0 + q);
      }
      // the arrays' last partial quad (last tile only): real entries element-wise
      if (ragged_end) {
        for (int li = bulk + lane; li < nslots; li += 32) {
          const int64_t k = kq0 + li;
          if (k < nnz_end) {
            cp_async_small<sizeof(I)>(smem_u32(data) + uint32_t(li) * sizeof(I), colind + k);
            if (!has_perm)
              cp_async_small<sizeof(T)>(smem_u32(data + off_b) + uint32_t(li) * sizeof(T),
                                        values + k);
            else
              cp_async_small<sizeof(O)>(smem_u32(data + off_b) + uint32_t(li) * sizeof(O),
                                        perm + k);
          }
        }
      }
      cp_async_arrive_noinc(full);

      if (++s == stages) {
        s = 0;
        phase ^= 1u;
      }
    }
    return;
  }

  // ========================= consumer warps ========================================
  int s = 0;
  uint32_t phase = 0;
  for (int64_t t = t_begin; t < t_end; ++t) {
    unsigned char* st = stage_base + size_t(s) * stage_bytes;
    unsigned char* data = st + kPipeHeaderBytes;
    mbar_wait(smem_u32(&bars[s]), phase);

    const PipeHeader* hp = reinterpret_cast<const PipeHeader*>(st);
    const int nr = hp->nr, lo = hp->lo, hi = hp->hi;
    const int uni = hp->uniform;
    const int64_t row0 = hp->row0;
    const O kq0 = O(hp->kq0);
    const I* col = reinterpret_cast<const I*>(data);
    T* prod = reinterpret_cast<T*>(data + hp->off_prod);
    const O* rowend = reinterpret_cast<const O*>(data + hp->off_rowend);

    // does a peer need rows of this tile?
    const bool scat = wants_rows(sc, row0, row0 + nr);
    auto put = [&](int64_t row, T v) {
      v = with_addend<ADD>(sc, row, v);
      y[row] = v;
      if (scat)
        scatter_store(sc, row, v);
    };

    int tb; // local index where the trailing partial row starts
    bool tail_from_prod;
    if (uni > 0 && nr > 0) {
      // ---- path 1: uniform tile ----------------------------------------------------
      const T* val = prod; // !has_perm: values are staged at off_prod == off_b
      const int e0 = int(rowend[0] - kq0);
      // first row: its head may lie in the previous tile and it may be long
      if (warp == 0) {
        const T sum = dot_warp<T, I>(col, val, x, lo, e0, lane);
        if (lane == 0)
          put(row0, alpha * sum);
      }
      T* yrow = y + row0 + 1;
      const int nrows = nr - 1;
      if (!scat)
        uniform_tile<kPipeConsumers, false, ADD, T, I>(uni, col, val, x, yrow, alpha, e0, nrows, tid, sc, row0 + 1);
      else
        uniform_tile<kPipeConsumers, true, ADD, T, I>(uni, col, val, x, yrow, alpha, e0, nrows, tid, sc, row0 + 1);
      tb = e0 + uni * nrows;
      tail_from_prod = false;
    } else {
      // ---- path 2 (A): flat products -------------------------------------------------
      const int nq = hp->nslots >> 2;
      if (!has_perm) {
        int q = tid;
        for (; q + kPipeConsumers < nq; q += 2 * kPipeConsumers)
          quad_products<T, I>(col, prod, prod, x, q, q + kPipeConsumers, true);
        if (q < nq)
          quad_products<T, I>(col, prod, prod, x, q, q, false);
      } else {
        const O* pst = reinterpret_cast<const O*>(data + hp->off_b);
        for (int q = tid; q < nq; q += kPipeConsumers)
          quad_products_perm<T, I, O>(col, pst, prod, values, x, q);
      }
      named_barrier_sync(1, kPipeConsumers);
      // ---- path 2 (B): rows out of shared memory ----------------------------------------
      const int nzt = hi - lo;
      if (nr > 0) {
        if (nzt <= nr * 12) {
          // short rows: one thread per row.  A row longer than kLongRow inside such
          // a tile is summed by the whole warp (ballot + broadcast of its bounds).
          for (int q0 = warp * 32; q0 < nr; q0 += kPipeConsumers) {
            const int q = q0 + lane;
            int b = 0, e = 0;
            if (q < nr) {
              b = q == 0 ? lo : int(rowend[q - 1] - kq0);
              e = int(rowend[q] - kq0);
            }
            const bool is_long = e - b > kLongRow;
            if (q < nr && !is_long)
              put(row0 + q, alpha * prod_sum_thread(prod, b, e));
            unsigned todo = __ballot_sync(0xffffffffu, is_long);
            while (todo) {
              const int src = __ffs(todo) - 1;
              todo &= todo - 1;
              const int bb = __shfl_sync(0xffffffffu, b, src);
              const int ee = __shfl_sync(0xffffffffu, e, src);
              const T sum = prod_sum_warp(prod, bb, ee, lane);
              if (lane == 0)
                put(row0 + q0 + src, alpha * sum);
            }
          }
        } else {
          // long rows: one warp per row
          for (int q = warp; q < nr; q += kPipeConsumerWarps) {
            const int b = q == 0 ? lo : int(rowend[q - 1] - kq0);
            const int e = int(rowend[q] - kq0);
            const T sum = prod_sum_warp(prod, b, e, lane);
            if (lane == 0)
              put(row0 + q, alpha * sum);
          }
        }
      }
      tb = nr > 0 ? int(rowend[nr - 1] - kq0) : lo;
      tail_from_prod = true;
    }

    // ---- trailing partial row -> carry --------------------------------------------------
    const int tlen = hi - tb;
    if (row0 + nr < rows && tlen > 0) {
      if (!tail_from_prod) {
        // uniform tile: the fragment is shorter than a row (at most 8 entries)
        if (warp == 1 % kPipeConsumerWarps) {
          const T sum = dot_warp<T, I>(col, prod, x, tb, hi, lane);
          if (lane == 0) {
            carry_row[t] = row0 + nr;
            carry_val[t] = sum;
          }
        }
      } else if (tlen <= 256) {
        if (warp == 0) {
          const T sum = prod_sum_warp(prod, tb, hi, lane);
          if (lane == 0) {
            carry_row[t] = row0 + nr;
            carry_val[t] = sum;
          }
        }
      } else {
        T sum = T(0);
        for (int i = tb + tid; i < hi; i += kPipeConsumers)
          sum += prod[i];
        sum = warp_reduce_sum(sum);
        if (lane == 0)
          s_red[warp] = sum;
        named_barrier_sync(2, kPipeConsumers);
        if (tid == 0) {
          T tot = T(0);
#pragma unroll
          for (int w = 0; w < kPipeConsumerWarps; ++w)
            tot += s_red[w];
          carry_row[t] = row0 + nr;
          carry_val[t] = tot;
        }
        named_barrier_sync(2, kPipeConsumers); // s_red reusable
      }
    } else if (tid == 0) {
      carry_row[t] = -1;
    }

    // release the stage
    __syncwarp();
    if (lane == 0)
      mbar_arrive(smem_u32(&bars[kPipeMaxStages + s]));
    if (++s == stages) {
      s = 0;
      phase ^= 1u;
    }
  }
}

// ============================================================================
// Warp-stream kernel: the general path for matrices bound by random gathers of x
// ============================================================================
// Such matrices (uniform random columns, R-MAT) are limited by the L1 tag stage —
// every gathered element of x is its own 128-byte line, one tag lookup per cycle per
// SM — so the kernel's only job is to keep gathers issuing at all times.  CTA-wide
// phases (load, barrier, reduce) leave that stage idle between phases; here every
// WARP is autonomous: it owns whole streams of the merged sequence (row ends ++
// nonzeros; a second, warp-granular merge-path table built by the inspect code),
// walks a stream in chunks of 256 nonzeros (two 128-bit loads of colind and of values
// per lane, eight gathers in flight per lane), and reduces the rows that end inside
// the chunk out of its own 256-entry slab of shared memory — one lane per row in
// storage order (the reference's order), the whole warp for rows longer than 32 — with
// nothing but __syncwarp().  A chunk in which no row ends (the inside of a hub row)
// never touches shared memory: the lanes' products go straight into a shuffle
// reduction.  Streams are dealt round-robin (warp w takes streams w, w + W, ...), so
// every warp samples the whole matrix and the load balances without atomics.  The
// stream's trailing partial row is a carry, added by the same fix-up kernel.
constexpr int kWsChunk = 256;   // nonzeros per warp step
constexpr int kWsWarps = 8;     // warps per CTA
// CTAs per SM = the register budget of the walk: 5 -> 48 registers, 4 -> 64.  With all of a
// chunk's loads issued together (kWsFlat) 40 registers (6 CTAs) spill; build-time switches for
// the A/B runs (profiles/r02_flat_walk_ab.jsonl).
#ifndef B200_WS_CTAS
#define B200_WS_CTAS 5
#endif
#ifndef B200_WS_CTAS_WIDE
#define B200_WS_CTAS_WIDE 4
#endif
constexpr int kWsCtasPerSm = B200_WS_CTAS;          // 4-byte values and indices
constexpr int kWsCtasPerSmWide = B200_WS_CTAS_WIDE; // 8-byte values or indices

template <typename T, typename I>
constexpr int ws_ctas_per_sm() {
  return (sizeof(T) == 8 || sizeof(I) == 8) ? kWsCtasPerSmWide : kWsCtasPerSm;
}

// The walk of one warp over its streams, shared by the two kernels below.  HUB: `colind`
// is the plan's re-encoded copy (hub.cu) in which a reference to hub column number s reads
// ~s (negative), and `hub` is the shared-window address of this CTA's copy of x at the
// hub columns.
template <typename T>
__device__ __forceinline__ T ld_hub(uint32_t hub, int slot) {
  // (a 32-bit shared address kept in one register: through a generic pointer the
  // compiler rebuilds the shared window's base under every predicate)
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(hub + uint32_t(slot) * 4u));
    return *reinterpret_cast<T*>(&r);
  } else {
    static_assert(sizeof(T) == 8, "4- or 8-byte element expected");
    unsigned long long r;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(hub + uint32_t(slot) * 8u));
    return *reinterpret_cast<T*>(&r);
  }
}

// CG: the gathers of x that do go to memory bypass L1 (ld.global.cg).  Measured on R-MAT scale
// 24 (profiles/r02_hub_ab_l1_bypass.jsonl): beside the hub table, where L1 is small and its hit
// rate under 1 %, bypassing wins (fp32 1.032 -> 1.021 ms, fp64 1.439 -> 1.360 ms) — the hub
// kernel always does it; in the plain walk, whose L1 is large, it loses badly (1.127 ->
// 1.691 ms) — the plain walk never does.
// (Tried and removed: loading the next chunk's indices one step ahead, +8 registers —
// 1.034 -> 1.032 ms on R-MAT scale 24, profiles/r01_hub_ab_rmat.jsonl.)
// L2 policy of the walk (build-time switches, measured in profiles/r02_l2_policy.txt):
// the streams of A (colind, values, permutation, row ends) are read once per product and
// carry an evict_first hint, so that on a matrix whose A is many times L2 (C5: 3.2 GB per
// GPU against 126 MB) they do not push out the columns of x that are gathered again and
// again; the gathers of x themselves may carry evict_last.
#ifndef B200_WS_A_EF
#define B200_WS_A_EF 0
#endif
#ifndef B200_WS_X_EL
#define B200_WS_X_EL 0
#endif
constexpr bool kWsStreamEF = B200_WS_A_EF != 0;
// the global-hub walk (x larger than L2): the compact table is the one operand worth keeping
// in L2 — its loads carry evict_last, the streams of A and the cold gathers evict_first
#ifndef B200_HUBG_L2
#define B200_HUBG_L2 0
#endif
constexpr bool kHubgL2 = B200_HUBG_L2 != 0;

template <bool CG, typename T>
__device__ __forceinline__ T ws_gather(const T* p) {
  if constexpr (CG)
    return __ldcg(p);
  else if constexpr (B200_WS_X_EL != 0)
    return ld_ro_el(p);
  else
    return ld_ro(p);
}

// one gather of x for the walk.  HUB 0: through L1; HUB 1: a negative index reads the
// shared-memory table, the rest bypass L1; HUB 2: a negative index reads the compact table
// in global memory.  L1 policy of HUB 2 (build-time switch B200_HUBG_L1, measured on the
// scale-27 shard, profiles/r02_hubg_l1_priorities.jsonl): 0 one plain load either way
// 1.92 ms; 1 table lines evict_last 1.88 ms; 2 (default) table evict_last AND the cold
// gathers of x not allocated in L1 — they come from DRAM and would only push the table's
// lines out — 1.69 ms; 3 cold gathers evict_first 1.73 ms.
template <int HUB, typename T, typename I>
__device__ __forceinline__ T ws_gather_x(const T* __restrict__ x, const T* __restrict__ xh,
                                         const uint32_t hub, const I c) {
  if constexpr (HUB == 1) {
    return c < I(0) ? ld_hub<T>(hub, int(~c)) : ws_gather<true>(x + c);
  } else if constexpr (HUB == 2) {
    const bool h = c < I(0);
#if defined(B200_HUBG_L1) && B200_HUBG_L1 == 0
    return ld_ro((h ? xh : x) + (h ? ~c : c));
#elif defined(B200_HUBG_L1) && B200_HUBG_L1 == 1
    return h ? ld_ro_l1_evict_last(xh + ~c) : ld_ro(x + c);
#elif defined(B200_HUBG_L1) && B200_HUBG_L1 == 3
    return h ? ld_ro(xh + ~c) : ld_ro_l1_evict_first(x + c);
#else
    if constexpr (kHubgL2)
      return h ? ld_ro_keep(xh + ~c) : ld_stream_ef(x + c);
    else
      return h ? ld_ro_l1_evict_last(xh + ~c) : ld_stream(x + c);
#endif
  } else {
    return ws_gather<false>(x + c);
  }
}

// B200_WS_FLAT=0 builds the walk without the branch-free load phase (A/B runs only)
#ifndef B200_WS_FLAT
#define B200_WS_FLAT 1
#endif
constexpr bool kWsFlat = B200_WS_FLAT != 0;

template <typename T, typename I, typename O, int HUB, int WARPS, bool ADD>
__device__ __forceinline__ void
ws_walk_streams(const O* __restrict__ rowptr, const I* __restrict__ colind,
                const T* __restrict__ values, const O* __restrict__ perm,
                const T* __restrict__ x, T* __restrict__ y, const T alpha,
                const int64_t* __restrict__ starts, const int64_t stream_first,
                const int64_t num_streams, const int64_t rows, const int64_t nnz_end,
                int64_t* __restrict__ carry_row, T* __restrict__ carry_val,
                const ScatterArgs<T>& sc, const int lane, const int warp, T* slab,
                const uint32_t hub, const T* __restrict__ xh = nullptr) {
  const int64_t gw = int64_t(blockIdx.x) * WARPS + warp;
  const int64_t nw = int64_t(gridDim.x) * WARPS;
  const bool has_perm = perm != nullptr;

  for (int64_t s = stream_first + gw; s < stream_first + num_streams; s += nw) {
    int64_t row = starts[2 * s];
    const int64_t k_s = starts[2 * s + 1];
    const int64_t row_e = starts[2 * s + 2];
    const bool scat = wants_rows(sc, row, row_e);
    auto put = [&](int64_t r, T v) {
      v = with_addend<ADD>(sc, r, v);
      y[r] = v;
      if (scat)
        scatter_store(sc, r, v);
    };
    // positions inside the stream are ints relative to `base`, the 16-byte aligned
    // origin of the first chunk
    const int64_t base = k_s & ~int64_t(3);
    const I* __restrict__ ci = colind + base;
    const T* __restrict__ va = values + base;
    const int k_e = int(starts[2 * s + 3] - base);
    const int64_t left = nnz_end - base;
    const int arr_end = left < int64_t(0x7fffffff) ? int(left) : 0x7fffffff;
    int rows_left = int(row_e - row);
    int cur = int(k_s - base); // where the still open row's part inside this stream begins
    T carry = T(0);            // lane 0: that part's sum over the chunks already done
    int k = cur;
    int kb = 0;
    do {
      const int kend = kb + kWsChunk < k_e ? kb + kWsChunk : k_e;
      // ---- products of this chunk: two quads per lane -----------------------------
      constexpr bool kEF = kWsStreamEF || (HUB == 2 && kHubgL2);
      T p[2][4];
      int re = 0x7fffffff;
      if (kWsFlat && kb + kWsChunk <= arr_end) {
        // Every quad of the chunk lies inside the arrays (all chunks but the arrays' last):
        // the loads need no bounds, so ALL of them — both quads of colind and of values, the
        // row ends, then the eight gathers — are issued from one basic block before anything
        // waits.  Written with a branch around each quad (the general form below) the
        // compiler keeps the quads apart: colind -> gathers -> products of quad 0, only then
        // the loads of quad 1, then the row ends: five memory latencies in a row per chunk
        // and four gathers in flight per lane instead of two latencies and eight.  A lane
        // whose quad lies outside [k, kend) reads the chunk's first quad instead (one address
        // for all such lanes: a broadcast, no extra sectors) and its products are zeroed.
        Quad<I> c[2];
        Quad<T> v[2];
        int kq[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = kb + 4 * (lane + 32 * u);
          kq[u] = (kk < kend && kk + 4 > k) ? kk : kb;
          c[u] = ld_stream_quad_p<kEF>(ci + kq[u]);
        }
        if (lane < rows_left)
          re = int(int64_t(ld_stream_p<kEF>(rowptr + row + 1 + lane)) - base);
        if (!has_perm) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
            v[u] = ld_stream_quad_p<kEF>(va + kq[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const Quad<O> pi = ld_stream_quad_p<kEF>(perm + base + kq[u]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              v[u].v[j] = ld_ro(values + pi.v[j]);
          }
        }
        T xv[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            xv[u][j] = ws_gather_x<HUB, T, I>(x, xh, hub, c[u].v[j]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = kb + 4 * (lane + 32 * u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            p[u][j] = (kk + j >= k && kk + j < kend) ? v[u].v[j] * xv[u][j] : T(0);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int kk = kb + 4 * (lane + 32 * u);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            p[u][j] = T(0);
          if (kk < kend && kk + 4 > k) {
            Quad<I> c;
            Quad<T> v;
            if (kk + 4 <= arr_end) {
              c = ld_stream_quad_p<kEF>(ci + kk);
              if (!has_perm) {
                v = ld_stream_quad_p<kEF>(va + kk);
              } else {
                const Quad<O> pi = ld_stream_quad_p<kEF>(perm + base + kk);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  v.v[j] = ld_ro(values + pi.v[j]);
              }
            } else { // the arrays' last partial quad
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool in = kk + j < arr_end;
                c.v[j] = in ? ld_stream_p<kEF>(ci + kk + j) : I(0);
                v.v[j] = !in ? T(0)
                             : (has_perm ? ld_ro(values + perm[base + kk + j])
                                         : ld_stream_p<kEF>(va + kk + j));
              }
            }
            T xv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              xv[j] = ws_gather_x<HUB, T, I>(x, xh, hub, c.v[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              p[u][j] = (kk + j >= k && kk + j < kend) ? v.v[j] * xv[j] : T(0);
          }
        }
        // ---- rows that end inside the chunk -----------------------------------------
        if (lane < rows_left)
          re = int(int64_t(ld_stream_p<kEF>(rowptr + row + 1 + lane)) - base);
      }
      unsigned mask = __ballot_sync(0xffffffffu, re <= kend);
      if (mask == 0u) {
        // the chunk lies inside one row: no shared memory, straight to the shuffle tree
        T sum = ((p[0][0] + p[0][1]) + (p[0][2] + p[0][3])) +
                ((p[1][0] + p[1][1]) + (p[1][2] + p[1][3]));
        sum = warp_reduce_sum(sum);
        carry += sum;
      } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          Vec4<T> q;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            q.v[j] = p[u][j];
          *reinterpret_cast<Vec4<T>*>(slab + 4 * (lane + 32 * u)) = q;
        }
        __syncwarp();
        for (;;) {
          const int nready = __popc(mask); // a prefix of the lanes: row ends ascend
          int b = __shfl_up_sync(0xffffffffu, re, 1);
          if (lane == 0)
            b = cur > kb ? cur : kb; // what lies before this chunk is in `carry`
          const bool mine = lane < nready;
          const int len = mine ? re - b : 0;
          if (mine && len <= 32) {
            // storage order, like the reference; loads four at a time so that a row
            // costs len/4 shared-memory round trips, not len
            const T* q = slab + (b - kb);
            T sum = T(0);
#pragma unroll 1
            for (int i = 0; i < len; i += 4) {
              const T t0 = q[i];
              const T t1 = i + 1 < len ? q[i + 1] : T(0);
              const T t2 = i + 2 < len ? q[i + 2] : T(0);
              const T t3 = i + 3 < len ? q[i + 3] : T(0);
              sum += t0;
              sum += t1;
              sum += t2;
              sum += t3;
            }
            if (lane == 0)
              sum += carry;
            put(row + lane, alpha * sum);
          }
          unsigned todo = __ballot_sync(0xffffffffu, mine && len > 32);
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int bb = __shfl_sync(0xffffffffu, b, src) - kb;
            const int ee = __shfl_sync(0xffffffffu, re, src) - kb;
            T sum = T(0);
            for (int i = bb + lane; i < ee; i += 32)
              sum += slab[i];
            sum = warp_reduce_sum(sum);
            if (lane == 0) {
              if (src == 0)
                sum += carry;
              put(row + src, alpha * sum);
            }
          }
          cur = __shfl_sync(0xffffffffu, re, nready - 1);
          row += nready;
          rows_left -= nready;
          carry = T(0);
          if (nready < 32 || rows_left <= 0)
            break;
          re = 0x7fffffff;
          if (lane < rows_left)
            re = int(int64_t(ld_stream_p<kWsStreamEF || (HUB == 2 && kHubgL2)>(rowptr + row + 1 + lane)) - base);
          mask = __ballot_sync(0xffffffffu, re <= kend);
          if (mask == 0u)
            break;
        }
        // what follows the last row end belongs to the row still open
        if (cur < kend) {
          T sum = T(0);
          for (int i = cur - kb + lane; i < kend - kb; i += 32)
            sum += slab[i];
          carry = warp_reduce_sum(sum);
        }
        __syncwarp(); // the slab is rewritten by the next chunk
      }
      k = kend;
      kb += kWsChunk;
    } while (k < k_e);
    if (lane == 0) {
      if (row_e < rows && cur < k_e) {
        carry_row[s] = row_e;
        carry_val[s] = carry;
      } else {
        carry_row[s] = -1;
      }
    }
  }
}

template <typename T, typename I, typename O, bool ADD>
__global__ void __launch_bounds__(kWsWarps * 32, ws_ctas_per_sm<T, I>())
spmv_warp_stream_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                        const T* __restrict__ values, const O* __restrict__ perm,
                        const T* __restrict__ x, T* __restrict__ y, const T alpha,
                        const int64_t* __restrict__ starts, const int64_t stream_first,
                        const int64_t num_streams, const int64_t rows,
                        const int64_t nnz_end,
                        int64_t* __restrict__ carry_row, T* __restrict__ carry_val,
                        const __grid_constant__ ScatterArgs<T> sc) {
  __shared__ __align__(16) T s_slab[kWsWarps][kWsChunk];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (gate_closed(sc))
    return;
  ws_walk_streams<T, I, O, 0, kWsWarps, ADD>(rowptr, colind, values, perm, x, y, alpha, starts,
                                            stream_first, num_streams, rows, nnz_end,
                                            carry_row, carry_val, sc, lane, warp, s_slab[warp],
                                            0u);
}

// ============================================================================
// Hub-stream kernel: the warp-stream walk with the most referenced columns of x held
// in shared memory
// ============================================================================
// On a matrix with skewed COLUMN popularity (R-MAT scale 24: the 32 K most referenced of
// 16.7 M columns take 37 % of the references) the warp-stream kernel sends almost every
// gather to L2 (L1 hit rate 11 % under a 2 GB stream).  Here the inspect phase (hub.cu)
// counts the references per column, picks the top H, and re-encodes a plan-owned copy
// of colind: a reference to hub number s is stored as ~s.  One CTA of 32 warps per SM
// loads x at the H hub columns into shared memory once (H loads per CTA and launch
// instead of one per reference), then runs the same walk; a negative index is a
// shared-memory read, everything else the same gather as before.  Same arithmetic in
// the same order as the warp-stream kernel: bit-identical y.  Measured:
// C4 1.13 -> 1.03 ms with 32768 columns; larger tables lose (L1 shrinks).
constexpr int kHubWarps = 32; // one CTA per SM

template <typename T, typename O, bool ADD>
__global__ void __launch_bounds__(kHubWarps * 32, 1)
spmv_hub_stream_kernel(const O* __restrict__ rowptr, const int32_t* __restrict__ hub_colind,
                       const T* __restrict__ values, const O* __restrict__ perm,
                       const T* __restrict__ x, T* __restrict__ y, const T alpha,
                       const int64_t* __restrict__ starts, const int64_t stream_first,
                       const int64_t num_streams, const int64_t rows,
                       const int64_t nnz_end,
                       int64_t* __restrict__ carry_row, T* __restrict__ carry_val,
                       const __grid_constant__ ScatterArgs<T> sc,
                       const int32_t* __restrict__ hub_cols, const int hub_n) {
  extern __shared__ __align__(16) unsigned char hub_smem[];
  T* slabs = reinterpret_cast<T*>(hub_smem);
  T* hub = slabs + kHubWarps * kWsChunk;
  for (int i = threadIdx.x; i < hub_n; i += kHubWarps * 32)
    hub[i] = ld_ro(x + ld_stream(hub_cols + i));
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t hub_addr = smem_u32(hub);
  asm volatile("" : "+r"(hub_addr)); // one register, not a recomputation at every use
  ws_walk_streams<T, int32_t, O, 1, kHubWarps, ADD>(
      rowptr, hub_colind, values, perm, x, y, alpha, starts, stream_first, num_streams, rows,
      nnz_end, carry_row, carry_val, sc, lane, warp, slabs + warp * kWsChunk, hub_addr);
}

// ============================================================================
// Hub table in GLOBAL memory: for x larger than L2
// ============================================================================
// When x does not fit in L2 (C5: R-MAT scale 27, x = 1.07 GB) a gather that misses L2 costs
// a 32-byte DRAM sector, and ncu shows almost all of them do (profiles/
// r02_ncu_c5shard_warp_stream.txt: 11.05 GB read for 3.3 GB of A — 0.9 sectors per stored
// entry): the popular columns are scattered over x, one useful element per 128-byte line,
// and do not survive in L2.  Here the inspect phase's column analysis (hub.cu) picks the
// columns referenced at least 3 times, up to half of L2 worth of them, hottest first, and
// re-encodes colind (hub number s -> ~s) exactly as for the shared-memory table; every
// product first gathers x at those columns into a compact table (hub_fill_kernel: H gathers
// instead of one per reference) and the walk then reads a negative index from the table:
// dense lines that stay in L2, the top of it in L1.  Same arithmetic in the same order as
// the warp-stream kernel: bit-identical y.  Occupancy and shape are the warp-stream kernel's.
template <typename T>
__global__ void __launch_bounds__(256)
hub_fill_kernel(const T* __restrict__ x, const int32_t* __restrict__ hub_cols, const int64_t h,
                T* __restrict__ xh) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < h; i += stride)
    xh[i] = ld_ro(x + ld_stream(hub_cols + i));
}

template <typename T, typename O, bool ADD>
__global__ void __launch_bounds__(kWsWarps * 32, ws_ctas_per_sm<T, int32_t>())
spmv_hubg_stream_kernel(const O* __restrict__ rowptr, const int32_t* __restrict__ hub_colind,
                        const T* __restrict__ values, const O* __restrict__ perm,
                        const T* __restrict__ x, T* __restrict__ y, const T alpha,
                        const int64_t* __restrict__ starts, const int64_t stream_first,
                        const int64_t num_streams, const int64_t rows,
                        const int64_t nnz_end,
                        int64_t* __restrict__ carry_row, T* __restrict__ carry_val,
                        const __grid_constant__ ScatterArgs<T> sc,
                        const T* __restrict__ xh) {
  __shared__ __align__(16) T s_slab[kWsWarps][kWsChunk];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  ws_walk_streams<T, int32_t, O, 2, kWsWarps, ADD>(rowptr, hub_colind, values, perm, x, y, alpha,
                                              starts, stream_first, num_streams, rows, nnz_end,
                                              carry_row, carry_val, sc, lane, warp, s_slab[warp],
                                              0u, xh);
}

// ============================================================================
// Fallback kernel: one tile per CTA
// ============================================================================
template <typename T, typename I, typename O, bool ADD>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_merge_tile_kernel(const O* __restrict__ rowptr,
                       const I* __restrict__ colind,
                       const T* __restrict__ values,
                       const O* __restrict__ perm, const T* __restrict__ x,
                       T* __restrict__ y, const T alpha,
                       const int64_t* __restrict__ tile_starts,
                       const int64_t tile_first, const int64_t rows,
                       const int64_t nnz_end, int64_t* __restrict__ carry_row,
                       T* __restrict__ carry_val, const int vec_ok,
                       const __grid_constant__ ScatterArgs<T> sc) {
  constexpr int THREADS = kSpmvThreads;
  constexpr int TILE = kSpmvMaxTileItems;
  constexpr int WARPS = THREADS / 32;
  constexpr int MAXLONG = TILE / kLongRow + 2;

  extern __shared__ __align__(16) unsigned char dyn_smem[];
  // s_prod[tile + slack] followed by s_rowend[tile + slack]; sized by the host
  __shared__ T s_red[WARPS];
  __shared__ int s_long[MAXLONG];
  __shared__ int s_nlong;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  if (gate_closed(sc))
    return;
  const int64_t t = tile_first + blockIdx.x;
  const int64_t row0 = tile_starts[2 * t], k0 = tile_starts[2 * t + 1];
  const int64_t row1 = tile_starts[2 * t + 2], k1 = tile_starts[2 * t + 3];
  const int nr = int(row1 - row0);
  const int64_t kq0 = k0 & ~int64_t(3); // aligned origin of the shared-memory index
  const int nslots = int(((k1 - kq0) + 3) & ~int64_t(3));
  T* s_prod = reinterpret_cast<T*>(dyn_smem);
  int* s_rowend = reinterpret_cast<int*>(dyn_smem + ((size_t(nslots + 4) * sizeof(T) + 15) & ~size_t(15)));
  const int nz_beg = int(k0 - kq0);
  const int nz_end = int(k1 - kq0);

  if (tid == 0)
    s_nlong = 0;
  const bool scat = wants_rows(sc, row0, row1);
  auto put = [&](int64_t row, T v) {
    v = with_addend<ADD>(sc, row, v);
    y[row] = v;
    if (scat)
      scatter_store(sc, row, v);
  };

  // ---- phase 1: row ends ----------------------------------------------------
  for (int q = tid; q < nr; q += THREADS)
    s_rowend[q] = int(int64_t(ld_stream(rowptr + row0 + 1 + q)) - kq0);

  // ---- phase 2: products, quad by quad over the aligned superset ----------------
  // A quad that lies inside the arrays is fetched with 128-bit streaming loads even
  // when it straddles the tile boundary (the neighbours' elements are masked out);
  // only the arrays' last partial quad, or unaligned arrays, take scalar loads.
  {
    const int nq = nslots >> 2;
    for (int q0 = 0; q0 < nq; q0 += 2 * THREADS) {
      Quad<I> c[2];
      Quad<T> v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + tid + u * THREADS;
        if (q < nq) {
          const int64_t k = kq0 + 4 * int64_t(q);
          if (vec_ok && k + 4 <= nnz_end) {
            c[u] = ld_stream_quad(colind + k);
            if (perm == nullptr) {
              v[u] = ld_stream_quad(values + k);
            } else {
              const Quad<O> pi = ld_stream_quad(perm + k);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                v[u].v[j] = (k + j >= k0 && k + j < k1) ? ld_ro(values + pi.v[j]) : T(0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool ok = k + j >= k0 && k + j < k1;
              c[u].v[j] = ok ? ld_stream(colind + k + j) : I(0);
              v[u].v[j] = !ok ? T(0)
                              : (perm == nullptr ? ld_stream(values + k + j)
                                                 : ld_ro(values + perm[k + j]));
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + tid + u * THREADS;
        if (q < nq) {
          const int64_t k = kq0 + 4 * int64_t(q);
          Vec4<T> p;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool ok = k + j >= k0 && k + j < k1;
            p.v[j] = ok ? v[u].v[j] * ld_ro(x + c[u].v[j]) : T(0);
          }
          *reinterpret_cast<Vec4<T>*>(&s_prod[k - kq0]) = p;
        }
      }
    }
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
          put(row0 + q, alpha * sum);
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
          put(row0 + q, alpha * sum);
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
          put(row0 + q, alpha * sum);
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
// summed in tile order by the thread of the run's LAST tile — the row itself ends in
// the tile after it, so a launch over tiles [fix_lo, fix_hi) completes exactly the
// rows that end in tiles [fix_lo + 1, fix_hi + 1): a chunk of tiles can be finished
// (and its rows shipped to the host) before later chunks have run.
//
// With a fused exchange this kernel is also where the iteration's cross-GPU barrier
// lives: the last CTA to finish tells every peer "my rows of step `epoch` are in your
// x" (a release store at system scope into its slot of the peer's flag array) and
// waits for the same word from every peer, so that when the stream moves on, the next
// x is complete here and this rank's old x is no longer being read anywhere.
template <typename T>
__global__ void __launch_bounds__(256)
spmv_carry_fixup_kernel(const int64_t* __restrict__ carry_row,
                        const T* __restrict__ carry_val, int64_t fix_lo,
                        int64_t fix_hi, int64_t num_tiles, T* __restrict__ y,
                        const T alpha, const __grid_constant__ ScatterArgs<T> sc,
                        const __grid_constant__ BarrierArgs bar) {
  if (gate_closed(sc))
    return; // (a gated call is never part of a fused exchange: no barrier to keep)
  const int64_t t = fix_lo + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  bool stored_to_peer = false;
  if (t < fix_hi) {
    const int64_t r = carry_row[t];
    if (r >= 0 && !(t + 1 < num_tiles && carry_row[t + 1] == r)) {
      int64_t s = t;
      while (s > 0 && carry_row[s - 1] == r)
        --s;
      T sum = carry_val[s];
      for (int64_t j = s + 1; j <= t; ++j)
        sum += carry_val[j];
      const T v = y[r] + alpha * sum;
      y[r] = v;
      if (sc.n > 0 && !sc.late) {
        scatter_store(sc, r, v);
        stored_to_peer = true;
      }
    }
  }
  if (bar.n == 0)
    return;
  // Order: (this kernel's own peer stores) -> count -> flag.  The product kernel's peer stores
  // are complete (kernel boundary on the stream); only a thread that stored to a peer HERE
  // needs the system-scope fence, everyone else orders through the block counter.
  __shared__ bool s_last;
  if (stored_to_peer)
    __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&bar.state[0], 1u);
    s_last = done == gridDim.x - 1;
    if (s_last)
      bar.state[0] = 0; // ready for the next step (stream order protects it)
  }
  __syncthreads();
  if (!s_last)
    return;
  if (sc.late) {
    // Late push: a few rows per peer (the halo of a banded matrix).  Stored from the product
    // kernel they made the one CTA that owns them a straggler — measured on C2 at 4 GPUs: peer
    // stores alone +13.5 us per step, the barrier alone +3.2 us (profiles/
    // r02_exchange_decomposition.txt) — so the product kernel does not store them at all, and
    // this block, which runs when every carry is in, copies them out of y (L2: other blocks
    // wrote them) in coalesced rows.
    __threadfence();
    constexpr int kBatch = 8; // loads of a batch in flight together, then its peer stores
    // (one index space over all destinations with 16 in flight measured no better: 0.2417 vs
    // 0.2396 ms per step at 4 GPUs)
    for (int d = 0; d < sc.n; ++d) {
      for (int64_t base = sc.lo[d]; base < sc.hi[d]; base += int64_t(blockDim.x) * kBatch) {
        T v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int64_t row = base + int64_t(u) * blockDim.x + threadIdx.x;
          v[u] = row < sc.hi[d] ? __ldcg(y + row) : T(0);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int64_t row = base + int64_t(u) * blockDim.x + threadIdx.x;
          if (row < sc.hi[d])
            sc.dst[d][row] = v[u];
        }
      }
    }
    __threadfence_system();
    __syncthreads();
  }
  if (int(threadIdx.x) >= bar.n)
    return;
  // (st.release.sys orders everything this thread observed — the block counter's chain — before
  // the flag: no separate system fence)
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(bar.remote[threadIdx.x]),
               "l"(bar.epoch)
               : "memory");
  const unsigned long long t0 = global_timer_ns();
  unsigned long long seen = 0;
  unsigned polls = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];"
                 : "=l"(seen)
                 : "l"(bar.local[threadIdx.x])
                 : "memory");
    if (seen >= bar.epoch)
      break;
    if ((++polls & 1023u) == 0 && global_timer_ns() - t0 > bar.timeout_ns) {
      // a peer is gone: do not hang the GPU.  The flag lives in host-mapped memory, where the
      // next execute on this plan reads it without a synchronisation and FAILS (cabi.cu).
      *reinterpret_cast<volatile unsigned int*>(bar.gave_up) = 1u;
      __threadfence_system();
      break;
    }
  }
}

// Units [T0, T1) of the active partition — tiles for the tile kernels, warp streams
// for the warp-stream kernel (the whole product: all of them).  A proper sub-range is
// one chunk of a host-buffer execute (host_exec.cu): chunks are launched in ascending
// order on one stream, and each completes the rows that end in its units.
template <typename T, typename I, typename O, bool ADD>
int launch_spmv(spblas_b200_plan* p, int variant, const void* alpha, const void* values,
                const void* x, void* y, int64_t T0, int64_t T1) {
  if (p->num_tiles == 0 && p->barrier.n == 0)
    return SPBLAS_B200_SUCCESS; // (a rank with no rows still takes part in the barrier)
  const int64_t ntiles = T1 - T0;
  const T a = *static_cast<const T*>(alpha);
  const bool perm = p->csr_perm != nullptr;
  const int vec_ok = spmv_vec_ok(p, values);
  const int64_t nnz_end = p->base + p->nnz;

  ScatterArgs<T> sc;
  sc.gate = p->gate;
  sc.gate_value = p->gate_value;
  sc.d = static_cast<const T*>(p->epi_d);
  sc.beta = sc.d ? *reinterpret_cast<const T*>(p->epi_beta) : T(0);
  sc.n = p->scatter.n;
  sc.multicast = p->scatter.multicast;
  sc.lo_min = INT64_MAX;
  sc.hi_max = INT64_MIN;
  for (int d = 0; d < kMaxPeers; ++d) {
    sc.dst[d] = d < sc.n ? static_cast<T*>(p->scatter.dst[d]) : nullptr;
    sc.lo[d] = d < sc.n ? p->scatter.lo[d] : 0;
    sc.hi[d] = d < sc.n ? p->scatter.hi[d] : 0;
    if (d < sc.n && sc.lo[d] < sc.hi[d]) {
      sc.lo_min = sc.lo[d] < sc.lo_min ? sc.lo[d] : sc.lo_min;
      sc.hi_max = sc.hi[d] > sc.hi_max ? sc.hi[d] : sc.hi_max;
    }
  }
  sc.late = 0;
  BarrierArgs bar;
  bar.n = p->barrier.n;
  bar.epoch = 0;
  bar.timeout_ns = p->barrier_timeout_ms * 1000000ull;
  bar.state = static_cast<unsigned int*>(p->barrier_state.p);
  bar.gave_up = p->barrier_gave_up_d;
  for (int d = 0; d < kMaxPeers; ++d) {
    bar.remote[d] = d < bar.n ? p->barrier.remote[d] : nullptr;
    bar.local[d] = d < bar.n ? p->barrier.local[d] : nullptr;
  }
  if (bar.n > 0)
    bar.epoch = ++p->barrier_epoch;
  // few rows to exchange (a halo), the whole product in one launch, a barrier to hang it on:
  // late push from the fix-up kernel; the product kernels then see no scatter at all
  ScatterArgs<T> sc_fix = sc;
  {
    int64_t push_rows = 0;
    for (int d = 0; d < sc.n; ++d)
      push_rows += sc.hi[d] > sc.lo[d] ? sc.hi[d] - sc.lo[d] : 0;
    const bool whole = T0 == 0 && T1 >= (variant == kVariantWarpStream || variant == kVariantHubStream ||
                                                 variant == kVariantHubGlobal
                                             ? p->ws_streams
                                             : p->num_tiles);
    if (sc.n > 0 && !sc.multicast && bar.n > 0 && whole && push_rows <= p->late_push_max_rows) {
      sc_fix.late = 1;
      sc.n = 0; // (lo_min / hi_max stay: wants_rows tests n first)
    }
  }

  cudaError_t e = cudaSuccess;
  // the carry arrays and the unit count of the active partition
  const bool hub = variant == kVariantHubStream;
  const bool hubg = variant == kVariantHubGlobal;
  const bool ws = variant == kVariantWarpStream || hub || hubg;
  const int64_t units = ws ? p->ws_streams : p->num_tiles;
  const int64_t* d_carry_row =
      static_cast<const int64_t*>(ws ? p->ws_carry_row.p : p->carry_row.p);
  const T* d_carry_val = static_cast<const T*>(ws ? p->ws_carry_val.p : p->carry_val.p);
  if (ntiles > 0 && hub) {
    if constexpr (sizeof(I) == 4) {
      // one CTA per SM: the walk's slabs and the hub table fill the SM's shared memory
      int64_t grid = (ntiles + kHubWarps - 1) / kHubWarps;
      if (grid > int64_t(p->num_sms))
        grid = p->num_sms;
      const size_t smem =
          (size_t(kHubWarps) * kWsChunk + size_t(p->hub_count)) * sizeof(T);
      auto kern = spmv_hub_stream_kernel<T, O, ADD>;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess)
        return cuda_fail(p, e, "cudaFuncSetAttribute(spmv_hub_stream_kernel)");
      // exactly what the slabs and the table need: the rest of the SM's array stays L1,
      // and every gather in flight holds an L1 line — with the maximum carve-out the
      // misses in flight bound the kernel (measured: slower than the plain walk with
      // 56 % of the gathers served from shared memory)
      int carve = p->ws_carveout;
      if (carve < 0) {
        carve = int(((smem + 1024) * 100 + p->smem_per_sm - 1) / p->smem_per_sm);
        carve = carve > 100 ? 100 : carve;
      }
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      // the encoded copy starts at the 16-byte aligned origin of the first stream
      // (entry base & ~3 of the caller's array); the kernel indexes it absolutely
      const int32_t* enc =
          static_cast<const int32_t*>(p->hub_colind.p) - (p->base & ~int64_t(3));
      kern<<<unsigned(grid), kHubWarps * 32, smem, p->stream>>>(
          static_cast<const O*>(p->csr_rowptr), enc, static_cast<const T*>(values),
          static_cast<const O*>(p->csr_perm), static_cast<const T*>(x), static_cast<T*>(y), a,
          static_cast<const int64_t*>(p->ws_starts.p), T0, ntiles, p->csr_rows, nnz_end,
          static_cast<int64_t*>(p->ws_carry_row.p), static_cast<T*>(p->ws_carry_val.p), sc,
          static_cast<const int32_t*>(p->hub_cols.p), int(p->hub_count));
      e = cudaGetLastError();
      if (e != cudaSuccess)
        return cuda_fail(p, e, "spmv_hub_stream_kernel");
    } else {
      return fail(p, SPBLAS_B200_NOT_SUPPORTED, "hub variant needs int32 column indices");
    }
  } else if (ntiles > 0 && hubg) {
    if constexpr (sizeof(I) == 4) {
      // x at the hub columns -> the compact table, then the walk
      const int64_t h = p->hub_count;
      if (h > 0) {
        const unsigned fgrid =
            unsigned(std::min<int64_t>((h + 255) / 256, int64_t(p->num_sms) * 8));
        hub_fill_kernel<T><<<fgrid, 256, 0, p->stream>>>(
            static_cast<const T*>(x), static_cast<const int32_t*>(p->hub_cols.p), h,
            static_cast<T*>(p->hub_x.p));
        e = cudaGetLastError();
        if (e != cudaSuccess)
          return cuda_fail(p, e, "hub_fill_kernel");
        p->last_launches += 1;
        p->total_launches += 1;
      }
      int64_t grid = (ntiles + kWsWarps - 1) / kWsWarps;
      if (grid > int64_t(p->num_sms) * ws_ctas_per_sm<T, I>())
        grid = int64_t(p->num_sms) * ws_ctas_per_sm<T, I>();
      int carve = p->ws_carveout;
      if (carve < 0) {
        const size_t need =
            size_t(ws_ctas_per_sm<T, I>()) * (kWsWarps * kWsChunk * sizeof(T) + 1024);
        carve = int((need * 100 + p->smem_per_sm - 1) / p->smem_per_sm);
        carve = carve > 100 ? 100 : carve;
      }
      auto kern = spmv_hubg_stream_kernel<T, O, ADD>;
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      const int32_t* enc =
          static_cast<const int32_t*>(p->hub_colind.p) - (p->base & ~int64_t(3));
      kern<<<unsigned(grid), kWsWarps * 32, 0, p->stream>>>(
          static_cast<const O*>(p->csr_rowptr), enc, static_cast<const T*>(values),
          static_cast<const O*>(p->csr_perm), static_cast<const T*>(x), static_cast<T*>(y), a,
          static_cast<const int64_t*>(p->ws_starts.p), T0, ntiles, p->csr_rows, nnz_end,
          static_cast<int64_t*>(p->ws_carry_row.p), static_cast<T*>(p->ws_carry_val.p), sc,
          static_cast<const T*>(p->hub_x.p));
      e = cudaGetLastError();
      if (e != cudaSuccess)
        return cuda_fail(p, e, "spmv_hubg_stream_kernel");
    } else {
      return fail(p, SPBLAS_B200_NOT_SUPPORTED, "hub variant needs int32 column indices");
    }
  } else if (ntiles > 0 && ws) {
    int64_t grid = (ntiles + kWsWarps - 1) / kWsWarps;
    if (grid > int64_t(p->num_sms) * ws_ctas_per_sm<T, I>())
      grid = int64_t(p->num_sms) * ws_ctas_per_sm<T, I>();
    // x is the only operand worth caching: L1 gets everything the slabs do not need
    // (the default carve-out is far larger and costs 9 % on R-MAT: fewer hub columns
    // of x stay in L1, and every L1 miss is a request cycle on the SM's L2 port)
    int carve = p->ws_carveout;
    if (carve < 0) {
      const size_t need =
          size_t(ws_ctas_per_sm<T, I>()) * (kWsWarps * kWsChunk * sizeof(T) + 1024);
      carve = int((need * 100 + p->smem_per_sm - 1) / p->smem_per_sm);
      carve = carve > 100 ? 100 : carve;
    }
    auto kern = spmv_warp_stream_kernel<T, I, O, ADD>;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    kern<<<unsigned(grid), kWsWarps * 32, 0, p->stream>>>(
        static_cast<const O*>(p->csr_rowptr), static_cast<const I*>(p->csr_colind),
        static_cast<const T*>(values), static_cast<const O*>(p->csr_perm),
        static_cast<const T*>(x), static_cast<T*>(y), a,
        static_cast<const int64_t*>(p->ws_starts.p), T0, ntiles, p->csr_rows, nnz_end,
        static_cast<int64_t*>(p->ws_carry_row.p), static_cast<T*>(p->ws_carry_val.p), sc);
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmv_warp_stream_kernel");
  } else if (ntiles <= 0) {
    // nothing to multiply: only the fix-up kernel's barrier runs
  } else if (variant == kVariantPipelined) {
    // Pipeline shape: stages x (header + tile data) of shared memory per CTA; shared
    // memory decides how many CTAs fit per SM.
    const int data_bytes = pipe_stage_data_bytes(p->tile_items, sizeof(T), sizeof(I),
                                                 sizeof(O), perm);
    const size_t per_stage = size_t(kPipeHeaderBytes) + size_t(data_bytes);
    const size_t budget = 227 * 1024;
    int ctas_per_sm = p->ctas_per_sm > 0 ? p->ctas_per_sm : 3;
    int stages = p->stages > 0 ? p->stages : 3;
    if (stages > kPipeMaxStages)
      stages = kPipeMaxStages;
    if (stages < 2)
      stages = 2;
    while (ctas_per_sm > 1 && (128 + 3 * per_stage + 1024) * ctas_per_sm > budget)
      --ctas_per_sm; // never trade the third stage for occupancy
    while (stages > 2 && (128 + stages * per_stage + 1024) * ctas_per_sm > budget)
      --stages;
    const size_t smem = 128 + size_t(stages) * per_stage;
    int64_t grid = int64_t(p->num_sms) * ctas_per_sm;
    if (grid > ntiles)
      grid = ntiles;
    auto launch = [&](auto kern, int threads) -> cudaError_t {
      cudaError_t e2 = cudaFuncSetAttribute(
          kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e2 != cudaSuccess)
        return e2;
      kern<<<unsigned(grid), threads, smem, p->stream>>>(
          static_cast<const O*>(p->csr_rowptr), static_cast<const I*>(p->csr_colind),
          static_cast<const T*>(values), static_cast<const O*>(p->csr_perm),
          static_cast<const T*>(x), static_cast<T*>(y), a,
          static_cast<const int64_t*>(p->tile_starts.p),
          static_cast<const int*>(p->tile_uniform.p), T0, ntiles, p->csr_rows, nnz_end,
          static_cast<int64_t*>(p->carry_row.p), static_cast<T*>(p->carry_val.p), stages,
          data_bytes, sc);
      return cudaGetLastError();
    };
    e = launch(spmv_pipe_kernel<T, I, O, 8, ADD>, 8 * 32 + 32);
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmv_pipe_kernel");
  } else {
    const size_t cap = size_t(p->tile_items) + kTileSlack;
    const size_t smem = ((cap * sizeof(T) + 15) & ~size_t(15)) + cap * sizeof(int) + 64;
    auto kern = spmv_merge_tile_kernel<T, I, O, ADD>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess)
      return cuda_fail(p, e, "cudaFuncSetAttribute(spmv_merge_tile_kernel)");
    kern<<<unsigned(ntiles), kSpmvThreads, smem, p->stream>>>(
        static_cast<const O*>(p->csr_rowptr), static_cast<const I*>(p->csr_colind),
        static_cast<const T*>(values), static_cast<const O*>(p->csr_perm),
        static_cast<const T*>(x), static_cast<T*>(y), a,
        static_cast<const int64_t*>(p->tile_starts.p), T0, p->csr_rows, nnz_end,
        static_cast<int64_t*>(p->carry_row.p), static_cast<T*>(p->carry_val.p), vec_ok, sc);
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmv_merge_tile_kernel");
  }
  // carries whose row ends inside [T0, T1): runs whose last tile is in [T0 - 1, T1 - 1)
  // (the partition's last tile never carries, so the final chunk simply runs to T1)
  const int64_t fix_lo = T0 > 0 ? T0 - 1 : 0;
  const int64_t fix_hi = T1 >= units ? units : T1 - 1;
  const int64_t fix_n = fix_hi > fix_lo ? fix_hi - fix_lo : 0;
  const unsigned fgrid = fix_n > 0 ? unsigned((fix_n + 255) / 256) : 1u;
  if (fix_n > 0 || bar.n > 0) {
    spmv_carry_fixup_kernel<T><<<fgrid, 256, 0, p->stream>>>(
        d_carry_row, d_carry_val, fix_lo, fix_hi, units, static_cast<T*>(y), a, sc_fix, bar);
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmv_carry_fixup_kernel");
  }
  const int launched = (ntiles > 0 ? 1 : 0) + ((fix_n > 0 || bar.n > 0) ? 1 : 0);
  p->last_launches += launched;
  p->total_launches += launched;
  return SPBLAS_B200_SUCCESS;
}

template <typename T, bool ADD>
int dispatch_index(spblas_b200_plan* p, int variant, const void* alpha, const void* values,
                   const void* x, void* y, int64_t T0, int64_t T1) {
  const bool i64 = p->idx_type == SPBLAS_B200_I64;
  const bool o64 = p->off_type == SPBLAS_B200_I64;
  if (!i64 && !o64)
    return launch_spmv<T, int32_t, int32_t, ADD>(p, variant, alpha, values, x, y, T0, T1);
  if (!i64 && o64)
    return launch_spmv<T, int32_t, int64_t, ADD>(p, variant, alpha, values, x, y, T0, T1);
  if (i64 && !o64)
    return launch_spmv<T, int64_t, int32_t, ADD>(p, variant, alpha, values, x, y, T0, T1);
  return launch_spmv<T, int64_t, int64_t, ADD>(p, variant, alpha, values, x, y, T0, T1);
}

} // namespace

} // namespace b200
