// device_utils.cuh — load/store and reduction primitives for the sm_100a kernels.
//
// The sparse-times-dense path is an HBM/L2-bound gather, so the
// primitives here are about memory behaviour, not math:
//   * 128-bit streaming loads of colind/values that do not allocate in L1
//     (ld.global.nc.L1::no_allocate) — A is touched exactly once per product, and
//     L1 is kept for the gathered dense operand;
//   * plain read-only loads (ld.global.nc) for the gathered operand x / B so that
//     neighbouring gathers hit in L1;
//   * streaming 128-bit stores for C.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace b200 {

__device__ __forceinline__ uint4 ld_stream_16(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}

__device__ __forceinline__ uint4 ld_ro_16(const void* p) {
  uint4 r;
  asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}

__device__ __forceinline__ void st_stream_16(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// streaming scalar load (no L1 allocation) for 4- and 8-byte elements
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return *reinterpret_cast<T*>(&r);
  } else {
    static_assert(sizeof(T) == 8, "4- or 8-byte element expected");
    unsigned long long r;
    asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return *reinterpret_cast<T*>(&r);
  }
}

// Four consecutive elements starting at a 16-byte aligned address, loaded with
// 128-bit streaming loads (one for 4-byte elements, two for 8-byte elements).
template <typename T>
struct Quad {
  T v[4];
};

template <typename T>
__device__ __forceinline__ Quad<T> ld_stream_quad(const T* p) {
  Quad<T> q;
  if constexpr (sizeof(T) == 4) {
    uint4 r = ld_stream_16(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      q.v[i] = *reinterpret_cast<const T*>(&w[i]);
  } else {
    static_assert(sizeof(T) == 8, "4- or 8-byte element expected");
    uint4 r0 = ld_stream_16(p);
    uint4 r1 = ld_stream_16(p + 2);
    const uint64_t w[4] = {
        (uint64_t(r0.y) << 32) | r0.x, (uint64_t(r0.w) << 32) | r0.z,
        (uint64_t(r1.y) << 32) | r1.x, (uint64_t(r1.w) << 32) | r1.z};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      q.v[i] = *reinterpret_cast<const T*>(&w[i]);
  }
  return q;
}

template <typename T>
__device__ __forceinline__ T ld_ro(const T* p) {
  return __ldg(p);
}

// ---- the same loads with an L2 eviction hint ---------------------------------------------
// createpolicy with fraction 1.0 folds to a constant in a uniform register (checked in the
// SASS: `LDG.E... desc[URx]`), so the hint costs no per-thread register.
//   evict_first: operands streamed once per product (colind, values, rowptr) — they leave
//                L2 first, so that they do not push out the dense operand;
//   evict_last:  the gathered dense operand (x, rows of B): what L2 is for.
#define B200_POLICY_EF "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n"
#define B200_POLICY_EL "createpolicy.fractional.L2::evict_last.b64 pol, 1.0;\n"

__device__ __forceinline__ uint4 ld_stream_16_ef(const void* p) {
  uint4 r;
  asm("{\n.reg .b64 pol;\n" B200_POLICY_EF
      "ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], pol;\n}"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}

template <typename T>
__device__ __forceinline__ T ld_stream_ef(const T* p) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm("{\n.reg .b64 pol;\n" B200_POLICY_EF
        "ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], pol;\n}"
        : "=r"(r)
        : "l"(p));
    return *reinterpret_cast<T*>(&r);
  } else {
    static_assert(sizeof(T) == 8, "4- or 8-byte element expected");
    unsigned long long r;
    asm("{\n.reg .b64 pol;\n" B200_POLICY_EF
        "ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], pol;\n}"
        : "=l"(r)
        : "l"(p));
    return *reinterpret_cast<T*>(&r);
  }
}

// gather of the dense operand, kept in L2 (and allocated in L1 like ld_ro)
template <typename T>
__device__ __forceinline__ T ld_ro_el(const T* p) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm("{\n.reg .b64 pol;\n" B200_POLICY_EL
        "ld.global.nc.L2::cache_hint.u32 %0, [%1], pol;\n}"
        : "=r"(r)
        : "l"(p));
    return *reinterpret_cast<T*>(&r);
  } else {
    static_assert(sizeof(T) == 8, "4- or 8-byte element expected");
    unsigned long long r;
    asm("{\n.reg .b64 pol;\n" B200_POLICY_EL
        "ld.global.nc.L2::cache_hint.u64 %0, [%1], pol;\n}"
        : "=l"(r)
        : "l"(p));
    return *reinterpret_cast<T*>(&r);
  }
}

__device__ __forceinline__ uint4 ld_ro_16_el(const void* p) {
  uint4 r;
  asm("{\n.reg .b64 pol;\n" B200_POLICY_EL
      "ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], pol;\n}"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}

// L1 eviction priorities for gathers (experiments of round 2: the compact hub table kept in
// L1, the cold gathers out of it)
template <typename T>
__device__ __forceinline__ T ld_ro_l1_evict_last(const T* p) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return *reinterpret_cast<T*>(&r);
  } else {
    unsigned long long r;
    asm("ld.global.nc.L1::evict_last.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return *reinterpret_cast<T*>(&r);
  }
}

// ... kept in L1 AND in L2 (the compact hub table: the one operand worth pinning)
template <typename T>
__device__ __forceinline__ T ld_ro_keep(const T* p) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm("{\n.reg .b64 pol;\n" B200_POLICY_EL
        "ld.global.nc.L1::evict_last.L2::cache_hint.u32 %0, [%1], pol;\n}"
        : "=r"(r)
        : "l"(p));
    return *reinterpret_cast<T*>(&r);
  } else {
    unsigned long long r;
    asm("{\n.reg .b64 pol;\n" B200_POLICY_EL
        "ld.global.nc.L1::evict_last.L2::cache_hint.u64 %0, [%1], pol;\n}"
        : "=l"(r)
        : "l"(p));
    return *reinterpret_cast<T*>(&r);
  }
}

template <typename T>
__device__ __forceinline__ T ld_ro_l1_evict_first(const T* p) {
  if constexpr (sizeof(T) == 4) {
    uint32_t r;
    asm("ld.global.nc.L1::evict_first.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return *reinterpret_cast<T*>(&r);
  } else {
    unsigned long long r;
    asm("ld.global.nc.L1::evict_first.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return *reinterpret_cast<T*>(&r);
  }
}

// EF = true: the streamed-once flavour with the evict_first hint
template <bool EF, typename T>
__device__ __forceinline__ Quad<T> ld_stream_quad_p(const T* p) {
  if constexpr (!EF) {
    return ld_stream_quad(p);
  } else {
    Quad<T> q;
    if constexpr (sizeof(T) == 4) {
      uint4 r = ld_stream_16_ef(p);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        q.v[i] = *reinterpret_cast<const T*>(&w[i]);
    } else {
      uint4 r0 = ld_stream_16_ef(p);
      uint4 r1 = ld_stream_16_ef(p + 2);
      const uint64_t w[4] = {
          (uint64_t(r0.y) << 32) | r0.x, (uint64_t(r0.w) << 32) | r0.z,
          (uint64_t(r1.y) << 32) | r1.x, (uint64_t(r1.w) << 32) | r1.z};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        q.v[i] = *reinterpret_cast<const T*>(&w[i]);
    }
    return q;
  }
}

template <bool EF, typename T>
__device__ __forceinline__ T ld_stream_p(const T* p) {
  if constexpr (EF)
    return ld_stream_ef(p);
  else
    return ld_stream(p);
}

// ---- warp reductions (warp-shuffle; sub-warp groups of `W` lanes) ------------
template <int W, typename T>
__device__ __forceinline__ T group_reduce_sum(T v, unsigned mask = 0xffffffffu) {
#pragma unroll
  for (int off = W / 2; off > 0; off >>= 1)
    v += __shfl_down_sync(mask, v, off, W);
  return v;
}

template <typename T>
__device__ __forceinline__ T warp_reduce_sum(T v) {
  return group_reduce_sum<32>(v);
}

} // namespace b200
