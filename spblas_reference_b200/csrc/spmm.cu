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

// the addend D of C = alpha A B + beta D: plain (coherent) loads — D may alias C
template <typename T, int VEC>
__device__ __forceinline__ BVec<T, VEC> load_d(const T* p) {
  BVec<T, VEC> r;
  if constexpr (VEC == 1) {
    r.v[0] = *p;
  } else {
    *reinterpret_cast<uint4*>(&r.v[0]) = *reinterpret_cast<const uint4*>(p);
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
                const int64_t k, const int64_t seg_limit, const T* D, const int64_t ldd,
                const T beta) {
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
        if (D != nullptr) { // C = alpha A B + beta D (D may alias C: read, then written, here)
          const BVec<T, VEC> dv = load_d<T, VEC>(D + row * ldd + c0);
#pragma unroll
          for (int u = 0; u < VEC; ++u)
            out.v[u] += beta * dv.v[u];
        }
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
                    const int64_t k, const T* D, const int64_t ldd, const T beta) {
  const int64_t s = blockIdx.x;
  const int64_t row = segments[3 * s];
  if (s > 0 && segments[3 * (s - 1)] == row)
    return; // not the first segment of its row
  for (int64_t c = threadIdx.x; c < k; c += blockDim.x) {
    T sum = T(0);
    for (int64_t q = s; q < num_segments && segments[3 * q] == row; ++q)
      sum += partial[q * k + c];
    T out = alpha * sum;
    if (D != nullptr)
      out += beta * D[row * ldd + c];
    C[row * ldc + c] = out;
  }
}

// ============================================================================
// Stream kernel (default for rows of B of >= kRingMinRowBytes): one warp = one run of the merged
// sequence (row ends ++ nonzeros); rows of B travel global -> shared memory by
// cp.async into a per-warp ring, lane l owns LB bytes (VEC columns) of C's row.
//
// Why: the group kernel above issues a batch of B loads into registers, waits for
// all of them, does the FMAs and starts over at every row; its loads in flight
// average a third of their peak (ncu, C3 k=32: 4.2 TB/s of DRAM reads, 14 stall
// cycles on long_scoreboard per issue).  A register ring does not fix that: a warp's
// loads are tracked by a handful of COUNTING scoreboards, so waiting for the oldest
// load of a rolling ring waits for the youngest load on the same scoreboard (tried:
// 1.4 ms against the group kernel's 0.79 ms on C3 k=32).  cp.async groups are the
// hardware's FIFO for exactly this: `cp.async.wait_group N` returns when all but the
// N most recent groups have landed, so a warp keeps R rows of B (4-8 KB) in flight
// at all times, across row boundaries, with no registers tied up.
//
// Per 4 entries: one wait_group, 4 x (row-end compare, SHFL of the value, one LDS,
// VEC FMAs), then the copies of the 4 entries R ahead into the slots just freed
// (k = 32 fp32: ONE cp.async instruction moves four 128-byte rows, 8 lanes x 16 B
// each) and a commit.  When a row end is reached the lane's accumulators are the
// row's slice of C and leave as one coalesced streaming store.
//
// Work split: the inspect-phase merge path cuts the merged sequence into as many
// equal runs as there are resident warps, so a hub row of a power-law matrix is
// shared by many warps (no segment kernels on this path) and a run of empty rows
// costs what its C stores cost.  A run's trailing partial row goes to a carry row;
// the fix-up kernel adds carries in stream order (deterministic, no atomics).
//
// (colind, value) pairs are loaded 32 at a time, one per lane, two chunks ahead, and
// broadcast by shuffle; row ends are kept 32 per warp, one batch ahead.  B is copied
// with an L2 policy that keeps a fraction of its lines evict_last (as much of B as
// fits beside the streams), A and C are evict_first / streaming, so the part of B that
// can live in the 126 MB L2 is not flushed by operands that are touched once.
// ============================================================================
constexpr int kRingThreads = 256;
constexpr int kRingWarps = kRingThreads / 32;
constexpr int kRingMinRowBytes = 256; // shorter rows of B: the group kernel
constexpr int kRingGroup = 4; // entries per cp.async commit group

// LB = bytes of a B row one lane consumes (4, 8, 16); GR = cp.async granule (4, 8, 16)
template <int LB, int GR>
struct RingShape {
  static constexpr int TB = 32 * LB;               // bytes of a B row per column tile
  static constexpr int R = TB >= 512 ? 16 : 32;    // ring slots (rows of B in flight)
  static constexpr int NG = R / kRingGroup;        // commit groups in flight
  static constexpr int LPE = TB / GR;              // lanes that copy one entry
  static constexpr int EPI = 32 / LPE;             // entries per cp.async instruction
  static constexpr int WARP_BYTES = R * TB;
  static_assert(GR == 16 || GR == LB, "granule is 16 bytes or the lane's slice");
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t policy_fraction_evict_last(float fraction) {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;"
               : "=l"(p)
               : "f"(fraction));
  return p;
}

__device__ __forceinline__ uint64_t policy_all_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int GR>
__device__ __forceinline__ void cp_async_policy(uint32_t dst, const void* src, uint64_t pol) {
  if constexpr (GR == 16)
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst),
                 "l"(src), "l"(pol)
                 : "memory");
  else
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], %2, %3;" ::"r"(dst),
                 "l"(src), "n"(GR), "l"(pol)
                 : "memory");
}

__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// one element of A's arrays, streamed: no L1 allocation, first out of L2
template <typename T>
__device__ __forceinline__ T ld_stream_policy(const T* p, uint64_t pol) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;"
                 : "=r"(r)
                 : "l"(p), "l"(pol));
    return *reinterpret_cast<T*>(&r);
  } else {
    static_assert(sizeof(T) == 8, "4- or 8-byte element expected");
    unsigned long long r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;"
                 : "=l"(r)
                 : "l"(p), "l"(pol));
    return *reinterpret_cast<T*>(&r);
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void store_c_stream(T* p, const BVec<T, VEC>& r) {
  constexpr int BYTES = VEC * int(sizeof(T));
  if constexpr (BYTES == 4) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p),
                 "r"(*reinterpret_cast<const uint32_t*>(&r.v[0]))
                 : "memory");
  } else if constexpr (BYTES == 8) {
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p),
                 "l"(*reinterpret_cast<const unsigned long long*>(&r.v[0]))
                 : "memory");
  } else {
    st_stream_16(p, *reinterpret_cast<const uint4*>(&r.v[0]));
  }
}

template <typename T, int VEC>
__device__ __forceinline__ BVec<T, VEC> lds_slice(uint32_t addr) {
  BVec<T, VEC> r;
  constexpr int BYTES = VEC * int(sizeof(T));
  if constexpr (BYTES == 4) {
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr));
    *reinterpret_cast<uint32_t*>(&r.v[0]) = w;
  } else if constexpr (BYTES == 8) {
    unsigned long long w;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(w) : "r"(addr));
    *reinterpret_cast<unsigned long long*>(&r.v[0]) = w;
  } else {
    uint4 w;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                 : "r"(addr));
    *reinterpret_cast<uint4*>(&r.v[0]) = w;
  }
  return r;
}

template <typename T, typename I, typename O, int LB, int GR>
struct RingState {
  using Shape = RingShape<LB, GR>;
  static constexpr int VEC = LB / int(sizeof(T));
  static constexpr int R = Shape::R, TB = Shape::TB, EPI = Shape::EPI, LPE = Shape::LPE;
  // operands
  const O* rowptr;
  const I* colind;
  const T* values;
  const O* perm;
  const unsigned char* Bsrc; // B + this tile's first column + this lane's granule, as bytes
  T* Cl;                     // C + c0
  int64_t ldc;
  const T* Dl;               // D + c0, or nullptr (beta = 0)
  int64_t ldd;
  T beta;
  unsigned ldb_bytes;
  T alpha;
  int64_t rows;
  int64_t ks, ke; // this stream's nonzeros
  uint64_t pol_b, pol_a;
  int lane;
  bool active;    // this lane owns columns of C
  bool copies;    // this lane's granule lies inside the tile's columns
  uint32_t ring_ld; // shared address of slot 0 + this lane's slice (consumer side)
  uint32_t ring_cp; // shared address of slot (lane / LPE) + this lane's granule (copy side)
  T acc[VEC];
  bool pending;
  // rows: the current row, its end, and two batches of 32 row ends
  int64_t row, rbase, rend;
  O re_cur, re_nxt;
  int rl; // rend relative to the current chunk's first entry (clipped)
  // (colind, value) of the current and the next chunk: one pair per lane
  I c_cur, c_nxt;
  T v_cur, v_nxt;

  __device__ __forceinline__ O load_rowend(int64_t r) const {
    // rows past the matrix: an end no entry index can equal
    return r < rows ? rowptr[r + 1] : (sizeof(O) == 8 ? O(0x7fffffffffffffffLL) : O(0x7fffffff));
  }
  __device__ __forceinline__ void load_meta(int64_t kc, I& c, T& v) const {
    const int64_t idx = kc + lane;
    c = I(0);
    v = T(0);
    if (idx >= ks && idx < ke) {
      c = ld_stream_policy(colind + idx, pol_a);
      v = perm == nullptr ? ld_stream_policy(values + idx, pol_a) : ld_ro(values + perm[idx]);
    }
  }
  __device__ __forceinline__ int rel(int64_t kc) const {
    const int64_t d = rend - kc;
    return d > int64_t(1 << 30) ? (1 << 30) : int(d);
  }
  // the current row is complete: its slice of C leaves as one coalesced store
  __device__ __forceinline__ void flush(int64_t kc) {
    if (active) {
      BVec<T, VEC> out;
#pragma unroll
      for (int u = 0; u < VEC; ++u)
        out.v[u] = alpha * acc[u];
      if (Dl != nullptr) {
#pragma unroll
        for (int u = 0; u < VEC; ++u)
          out.v[u] += beta * Dl[row * ldd + u];
      }
      store_c_stream<T, VEC>(Cl + row * ldc, out);
    }
#pragma unroll
    for (int u = 0; u < VEC; ++u)
      acc[u] = T(0);
    pending = false;
    ++row;
    int off = int(row - rbase);
    if (off == 32) {
      re_cur = re_nxt;
      rbase += 32;
      re_nxt = load_rowend(rbase + 32 + lane);
      off = 0;
    }
    rend = int64_t(__shfl_sync(0xffffffffu, re_cur, off));
    rl = rel(kc);
  }

  // copies of the kRingGroup entries at chunk positions [q0, q0 + 4) (q0 may lie in the
  // next chunk: q0 >= 32), into the ring slots starting at byte `slot`; one commit group
  template <bool CHECK>
  __device__ __forceinline__ void issue_group(int q0, uint32_t slot, int lo, int hi) {
    const I c_src = q0 < 32 ? c_cur : c_nxt;
#pragma unroll
    for (int i = 0; i < kRingGroup / EPI; ++i) {
      const int q = q0 + i * EPI + lane / LPE; // this lane's entry
      const I col = __shfl_sync(0xffffffffu, c_src, q & 31);
      if (copies && (!CHECK || (q >= lo && q < hi)))
        cp_async_policy<GR>(ring_cp + slot + uint32_t(i * EPI * TB),
                            Bsrc + size_t(col) * size_t(ldb_bytes), pol_b);
    }
    cp_async_commit();
  }

  __device__ __forceinline__ void consume(int u, uint32_t slot_of_u) {
    const T a = __shfl_sync(0xffffffffu, v_cur, u);
    const BVec<T, VEC> b = lds_slice<T, VEC>(ring_ld + slot_of_u);
#pragma unroll
    for (int q = 0; q < VEC; ++q)
      acc[q] += a * b.v[q];
  }

  // one chunk of 32 entries, four at a time.  CHECK = false when the chunk and the R
  // entries after it lie inside the stream.  The group loop is NOT unrolled (ring slots
  // and shuffle sources are computed, not baked in): the body stays within the
  // instruction cache; a group with no row end inside takes a path with no per-entry
  // compare.
  template <bool CHECK>
  __device__ __forceinline__ void chunk(int64_t kc) {
    const int lo = CHECK ? (ks - kc > 64 ? 64 : int(ks - kc)) : 0;
    const int hi = CHECK ? (ke - kc > 64 ? 64 : int(ke - kc)) : 64;
#pragma unroll 1
    for (int g0 = 0; g0 < 32; g0 += kRingGroup) {
      const uint32_t slot = uint32_t(g0 % R) * uint32_t(TB);
      cp_async_wait<Shape::NG - 1>(); // the oldest group in flight: entries g0 .. g0+3
      __syncwarp();
      const bool plain = rl >= g0 + kRingGroup && (!CHECK || (g0 >= lo && g0 + kRingGroup <= hi));
      if (plain) {
#pragma unroll
        for (int j = 0; j < kRingGroup; ++j)
          consume(g0 + j, slot + uint32_t(j * TB));
        pending = true;
      } else {
#pragma unroll
        for (int j = 0; j < kRingGroup; ++j) {
          const int u = g0 + j;
          if (!CHECK || (u >= lo && u < hi)) {
            while (rl == u) // rows that end before entry u (empty rows: several)
              flush(kc);
            consume(u, slot + uint32_t(j * TB));
            pending = true;
          }
        }
      }
      __syncwarp(); // every lane has read the slots before they are refilled
      issue_group<CHECK>(g0 + R, slot, lo, hi);
    }
  }
};

template <typename T, typename I, typename O, int LB, int GR, int MINB>
__global__ void __launch_bounds__(kRingThreads, MINB)
spmm_ring_kernel(const O* __restrict__ rowptr, const I* __restrict__ colind,
                 const T* __restrict__ values, const O* __restrict__ perm,
                 const T* __restrict__ B, const unsigned ldb_bytes, T* __restrict__ C,
                 const int64_t ldc, const T alpha, const int64_t rows, const int64_t k,
                 const int64_t* __restrict__ starts, int64_t* __restrict__ carry_row,
                 T* __restrict__ carry_val, const float l2_fraction, const T* D,
                 const int64_t ldd, const T beta) {
  using State = RingState<T, I, O, LB, GR>;
  using Shape = RingShape<LB, GR>;
  constexpr int VEC = State::VEC;
  constexpr int R = Shape::R;
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t w = int64_t(blockIdx.x) * kRingWarps + warp;
  const int64_t tile_c0 = int64_t(blockIdx.y) * 32 * VEC; // first column of the tile
  const int64_t c0 = tile_c0 + int64_t(lane) * VEC;
  const int64_t row_s = starts[2 * w], row_e = starts[2 * w + 2];

  State s;
  s.ks = starts[2 * w + 1];
  s.ke = starts[2 * w + 3];
  if (row_s == row_e && s.ks == s.ke) { // an empty stream (more warps than work)
    if (blockIdx.y == 0 && lane == 0)
      carry_row[w] = -1;
    return;
  }
  s.rowptr = rowptr;
  s.colind = colind;
  s.values = values;
  s.perm = perm;
  s.active = c0 < k;
  s.Cl = C + (s.active ? c0 : 0);
  s.ldc = ldc;
  s.Dl = D != nullptr ? D + (s.active ? c0 : 0) : nullptr;
  s.ldd = ldd;
  s.beta = beta;
  s.ldb_bytes = ldb_bytes;
  s.alpha = alpha;
  s.rows = rows;
  s.lane = lane;
  {
    // copy side: lane -> (entry lane / LPE of an instruction, granule lane % LPE)
    const int gran = lane % Shape::LPE;
    const int64_t gcol = tile_c0 + int64_t(gran) * (GR / int(sizeof(T)));
    s.copies = gcol < k;
    s.Bsrc = reinterpret_cast<const unsigned char*>(B + (s.copies ? gcol : 0));
    const uint32_t ring = smem_addr(ring_smem) + uint32_t(warp) * Shape::WARP_BYTES;
    s.ring_ld = ring + uint32_t(lane) * LB;
    s.ring_cp = ring + uint32_t(lane / Shape::LPE) * Shape::TB + uint32_t(gran) * GR;
  }
  s.pol_b = policy_fraction_evict_last(l2_fraction);
  s.pol_a = policy_all_evict_first();
#pragma unroll
  for (int u = 0; u < VEC; ++u)
    s.acc[u] = T(0);
  s.pending = false;
  s.row = row_s;
  s.rbase = row_s;
  s.re_cur = s.load_rowend(s.rbase + lane);
  s.re_nxt = s.load_rowend(s.rbase + 32 + lane);
  s.rend = int64_t(__shfl_sync(0xffffffffu, s.re_cur, 0));

  int64_t kc = s.ks & ~int64_t(31);
  s.load_meta(kc, s.c_cur, s.v_cur);
  s.load_meta(kc + 32, s.c_nxt, s.v_nxt);
  {
    // prologue: the copies of the first R entries
    const int lo = int(s.ks - kc);
    const int hi = s.ke - kc > 64 ? 64 : int(s.ke - kc);
#pragma unroll
    for (int q0 = 0; q0 < R; q0 += kRingGroup)
      s.template issue_group<true>(q0, uint32_t(q0 * Shape::TB), lo, hi);
  }
  for (; kc < s.ke; kc += 32) {
    I c_nn;
    T v_nn;
    s.load_meta(kc + 64, c_nn, v_nn);
    s.rl = s.rel(kc);
    if (kc >= s.ks && kc + 32 + R <= s.ke)
      s.template chunk<false>(kc);
    else
      s.template chunk<true>(kc);
    s.c_cur = s.c_nxt;
    s.v_cur = s.v_nxt;
    s.c_nxt = c_nn;
    s.v_nxt = v_nn;
  }
  cp_async_wait<0>();
  // rows that end at the stream's last entry, and empty rows after it
  while (s.row < row_e)
    s.flush(kc);
  // the trailing partial row continues in the next stream: carry
  if (s.pending) {
    if (s.active) {
#pragma unroll
      for (int u = 0; u < VEC; ++u)
        carry_val[w * k + c0 + u] = s.acc[u];
    }
    if (blockIdx.y == 0 && lane == 0)
      carry_row[w] = s.row;
  } else if (blockIdx.y == 0 && lane == 0) {
    carry_row[w] = -1;
  }
}

// One CTA per stream: the first stream of a run carrying into the same row adds the
// run's partial rows, in stream order, to the row of C written by the stream that
// held the row's end.
template <typename T>
__global__ void __launch_bounds__(128)
spmm_carry_fixup_kernel(const int64_t* __restrict__ carry_row,
                        const T* __restrict__ carry_val, const int64_t streams,
                        T* __restrict__ C, const int64_t ldc, const T alpha,
                        const int64_t k) {
  const int64_t w = blockIdx.x;
  const int64_t r = carry_row[w];
  if (r < 0 || (w > 0 && carry_row[w - 1] == r))
    return;
  for (int64_t c = threadIdx.x; c < k; c += blockDim.x) {
    T sum = carry_val[w * k + c];
    for (int64_t j = w + 1; j < streams && carry_row[j] == r; ++j)
      sum += carry_val[j * k + c];
    C[r * ldc + c] += alpha * sum;
  }
}

template <typename T, typename I, typename O, int LB, int GR>
int launch_spmm_ring(spblas_b200_plan* p, const T alpha, const void* values,
                     const void* B, int64_t ldb, void* C, int64_t ldc, int64_t k) {
  using Shape = RingShape<LB, GR>;
  constexpr int VEC = LB / int(sizeof(T));
  constexpr int MINB = Shape::WARP_BYTES <= 4096 ? 4 : 3;
  constexpr size_t smem = size_t(kRingWarps) * Shape::WARP_BYTES;
  auto kern = spmm_ring_kernel<T, I, O, LB, GR, MINB>;
  cudaError_t e =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess)
    return cuda_fail(p, e, "cudaFuncSetAttribute(spmm_ring_kernel)");
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRingThreads, smem);
  if (e != cudaSuccess || per_sm < 1)
    per_sm = 1;
  if (p->spmm_ctas_per_sm > 0 && p->spmm_ctas_per_sm < per_sm)
    per_sm = p->spmm_ctas_per_sm;
  int64_t streams = int64_t(p->num_sms) * per_sm * kRingWarps;
  // small problems: no more streams than runs of 64 merge items
  const int64_t total = p->csr_rows + p->nnz;
  const int64_t cap = ((total + 63) / 64 + kRingWarps - 1) / kRingWarps * kRingWarps;
  if (streams > cap)
    streams = cap;
  if (p->spmm_streams != streams)
    if (int rc = build_stream_partition(p, streams))
      return rc;
  if (int rc = reserve(p, p->spmm_carry_val, size_t(streams) * size_t(k) * sizeof(T)))
    return rc;
  // share of B to keep in L2: what fits in ~3/4 of it (the rest serves the streams)
  float frac = p->spmm_l2_fraction;
  if (frac < 0.f) {
    const double b_bytes = double(p->csr_cols) * double(ldb) * sizeof(T);
    const double room = 0.75 * double(p->l2_bytes);
    frac = b_bytes <= room ? 1.f : float(room / b_bytes);
  }
  if (frac > 1.f)
    frac = 1.f;
  const int64_t col_tiles = (k + 32 * VEC - 1) / (32 * VEC);
  const dim3 grid{unsigned(streams / kRingWarps), unsigned(col_tiles), 1u};
  kern<<<grid, kRingThreads, smem, p->stream>>>(
      static_cast<const O*>(p->csr_rowptr), static_cast<const I*>(p->csr_colind),
      static_cast<const T*>(values), static_cast<const O*>(p->csr_perm),
      static_cast<const T*>(B), unsigned(ldb * int64_t(sizeof(T))), static_cast<T*>(C), ldc,
      alpha, p->csr_rows, k, static_cast<const int64_t*>(p->spmm_starts.p),
      static_cast<int64_t*>(p->spmm_carry_row.p), static_cast<T*>(p->spmm_carry_val.p), frac,
      static_cast<const T*>(p->epi_d), p->epi_ldd,
      p->epi_d ? *reinterpret_cast<const T*>(p->epi_beta) : T(0));
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return cuda_fail(p, e, "spmm_ring_kernel");
  spmm_carry_fixup_kernel<T><<<unsigned(streams), 128, 0, p->stream>>>(
      static_cast<const int64_t*>(p->spmm_carry_row.p),
      static_cast<const T*>(p->spmm_carry_val.p), streams, static_cast<T*>(C), ldc, alpha, k);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return cuda_fail(p, e, "spmm_carry_fixup_kernel");
  p->last_launches += 2;
  p->total_launches += 2;
  return SPBLAS_B200_SUCCESS;
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
          static_cast<const T*>(B), ldb, static_cast<T*>(C), ldc, alpha, rows, k, seg_limit,
          static_cast<const T*>(p->epi_d), p->epi_ldd,
          p->epi_d ? *reinterpret_cast<const T*>(p->epi_beta) : T(0));
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
        static_cast<const T*>(p->seg_partial.p), static_cast<T*>(C), ldc, alpha, k,
        static_cast<const T*>(p->epi_d), p->epi_ldd,
        p->epi_d ? *reinterpret_cast<const T*>(p->epi_beta) : T(0));
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return cuda_fail(p, e, "spmm_combine_kernel");
    launches += 2;
  }
  p->last_launches += launches;
  p->total_launches += launches;
  return SPBLAS_B200_SUCCESS;
}

// Does a fair share of the stored entries sit in long rows?  Lower bound from the inspect
// phase's log2 row-length histogram (bin b holds the rows of 2^(b-1) .. 2^b - 1 entries): the
// entries of rows with at least 256.  On such a matrix the group kernel — one group of lanes
// per row, rows of up to 4096 entries walked by a single group — leaves most lanes waiting for
// the groups that drew the long rows, and the stream kernel, whose merge-path runs cut rows
// wherever the work is, wins at EVERY width of B (R-MAT scale 22, profiles/
// r02_spmm_rmat_narrow.jsonl: fp32 k = 8 / 16 / 32 6.2 / 9.4 / 10.5 ms -> 2.9 / 2.9 / 2.8 ms,
// fp64 k = 4 / 16 6.8 / 11.9 -> 3.5 / 3.5 ms).
bool heavy_tailed_rows(const spblas_b200_plan* p) {
  if (!p->have_hist || p->nnz <= 0)
    return false;
  double in_long_rows = 0.0;
  for (int b = 9; b < SPBLAS_B200_HIST_BINS; ++b)
    in_long_rows += double(p->hist[b]) * double(1ull << (b - 1));
  return in_long_rows >= 0.10 * double(p->nnz);
}

// One pass: C[:, 0:k] = alpha * A * B[:, 0:k] for the k columns starting at B / C.
template <typename T, typename I, typename O>
int spmm_pass(spblas_b200_plan* p, const void* alpha, const void* values,
              const void* B, int64_t ldb, void* C, int64_t ldc, int64_t k) {
  constexpr int V = 16 / sizeof(T);
  const T a = *static_cast<const T*>(alpha);
  const auto aligned16 = [](const void* q) {
    return (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
  };
  const bool vec = (k % V == 0) && (ldb % V == 0) && (ldc % V == 0) &&
                   aligned16(B) && aligned16(C) &&
                   (p->epi_d == nullptr || (p->epi_ldd % V == 0 && aligned16(p->epi_d)));
  // Stream kernel when a row of B is at least kRingMinRowBytes long: there it reaches
  // the DRAM peak on its traffic (C3 k=128: 6.4 TB/s), while on 128-byte rows both
  // kernels sit at the same random-access DRAM ceiling (~4.3 TB/s) and the group kernel
  // spends fewer instructions — unless the row lengths are heavy-tailed (the inspect phase's
  // histogram decides: heavy_tailed_rows).  B's row pitch must fit 32 bits, <= 65535 column tiles.
  const bool ring_ok = ldb * int64_t(sizeof(T)) < (int64_t(1) << 31) && k <= 65535 * 32;
  const bool ring = p->spmm_forced >= 0
                        ? (p->spmm_forced == 1 && ring_ok)
                        : (ring_ok && (k * int64_t(sizeof(T)) >= kRingMinRowBytes ||
                                       heavy_tailed_rows(p)));
  if (ring) {
    constexpr int S = int(sizeof(T));
    // 16-byte granules need 16-byte aligned rows of B (C may be anywhere unless the
    // lanes store 16 bytes)
    const bool b16 = (k % V == 0) && (ldb % V == 0) && aligned16(B);
    if (vec && k > 32 * (8 / S)) { // wide C: 16 bytes per lane
      p->spmm_variant = 1000 + V;
      return launch_spmm_ring<T, I, O, 16, 16>(p, a, values, B, ldb, C, ldc, k);
    }
    p->spmm_variant = 1001;
    if (b16)
      return launch_spmm_ring<T, I, O, S, 16>(p, a, values, B, ldb, C, ldc, k);
    return launch_spmm_ring<T, I, O, S, S>(p, a, values, B, ldb, C, ldc, k);
  }
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

// (Column slicing — running the product as several passes over A, each against a column
// slice of B — was built and measured in round 2 and removed: L2 allocates whole 128-byte
// lines, so a 32- or 64-byte slice of a row-major B occupies as much of L2 as the full row
// and every pass costs what the whole product costs (C3 k=32: 0.78 ms in one pass, 1.55 ms
// in two, 3.07 ms in four; k=128: 2.55 -> 3.50 ms in four; profiles/
// r02_spmm_column_slicing_negative.jsonl).)
template <typename T, typename I, typename O>
int pick_shape(spblas_b200_plan* p, const void* alpha, const void* values,
               const void* B, int64_t ldb, void* C, int64_t ldc, int64_t k) {
  return spmm_pass<T, I, O>(p, alpha, values, B, ldb, C, ldc, k);
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
