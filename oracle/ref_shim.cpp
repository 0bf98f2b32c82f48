// ref_shim.cpp — exposes the UNMODIFIED reference (headers included where they lie
// under /root/reference/include; nothing is copied) as plain C functions, so that
//   * tests can pin oracle/spblas_oracle.c against the real spblas::multiply,
//   * tests/golden/make_golden.py can generate the committed golden vectors,
//   * bench.py can time the reference's own CPU path (cpu_baseline.kind = "reference").
// Built by oracle/Makefile into oracle/_ref/libspblas_ref.so (git-ignored, travels to
// the GPU box).  TEST INFRASTRUCTURE ONLY — never loaded by the product path.
//
// Compiled with hidden visibility: the reference CPU backend defines
// spblas::index_t = size_t and an operation_info_t without backend state
// (detail/types.hpp:28-31, detail/operation_info_t.hpp:28-104), which must never meet
// the B200 backend's definitions of the same names in one symbol namespace.
#include <numeric>
#include <stdexcept>

#include <spblas/spblas.hpp>

#include <cstdint>
#include <span>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

template <typename T, typename I, typename O>
spblas::csr_view<T, I, O> make_csr(int64_t m, int64_t n, int64_t nnz,
                                   const O* rowptr, const I* colind,
                                   const T* values) {
  return spblas::csr_view<T, I, O>(const_cast<T*>(values), const_cast<O*>(rowptr),
                                   const_cast<I*>(colind),
                                   spblas::index<I>(I(m), I(n)), O(nnz));
}

template <typename T, typename I, typename O>
spblas::csc_view<T, I, O> make_csc(int64_t m, int64_t n, int64_t nnz,
                                   const O* colptr, const I* rowind,
                                   const T* values) {
  return spblas::csc_view<T, I, O>(const_cast<T*>(values), const_cast<O*>(colptr),
                                   const_cast<I*>(rowind),
                                   spblas::index<I>(I(m), I(n)), O(nnz));
}

// multiply(a, x, y) with the scaled views the caller asked for; `inspect` selects
// the two-phase spelling multiply_inspect + multiply(info, ...)
// (examples/spmm_csr.cpp:45-46).
template <typename A, typename T>
void spmv(A a, int has_aa, T alpha_a, int has_ax, T alpha_x, const T* x,
          int64_t n, T* y, int64_t m, int inspect) {
  std::span<T> xs(const_cast<T*>(x), size_t(n));
  std::span<T> ys(y, size_t(m));
  auto run = [&](auto&& av, auto&& xv) {
    if (inspect) {
      auto info = spblas::multiply_inspect(av, xv, ys);
      spblas::multiply(info, av, xv, ys);
    } else {
      spblas::multiply(av, xv, ys);
    }
  };
  if (has_aa && has_ax)
    run(spblas::scaled(alpha_a, a), spblas::scaled(alpha_x, xs));
  else if (has_aa)
    run(spblas::scaled(alpha_a, a), xs);
  else if (has_ax)
    run(a, spblas::scaled(alpha_x, xs));
  else
    run(a, xs);
}

template <typename A, typename T, typename I>
void spmm(A a, int has_aa, T alpha_a, int has_ab, T alpha_b, const T* B,
          int64_t n, int64_t k, T* C, int64_t m, int inspect) {
  spblas::mdspan_row_major<T, I> b(const_cast<T*>(B), I(n), I(k));
  spblas::mdspan_row_major<T, I> c(C, I(m), I(k));
  auto run = [&](auto&& av, auto&& bv) {
    if (inspect) {
      auto info = spblas::multiply_inspect(av, bv, c);
      spblas::multiply(info, av, bv, c);
    } else {
      spblas::multiply(av, bv, c);
    }
  };
  if (has_aa && has_ab)
    run(spblas::scaled(alpha_a, a), spblas::scaled(alpha_b, b));
  else if (has_aa)
    run(spblas::scaled(alpha_a, a), b);
  else if (has_ab)
    run(a, spblas::scaled(alpha_b, b));
  else
    run(a, b);
}

} // namespace

#define DEF_OPS(T, TN, I, IN, O, ON)                                               \
  REF_API void ref_csr_spmv_##TN##_##IN##_##ON(                                    \
      int64_t m, int64_t n, int64_t nnz, const O* rowptr, const I* colind,         \
      const T* values, int has_aa, T alpha_a, int has_ax, T alpha_x, const T* x,   \
      T* y, int inspect) {                                                         \
    spmv(make_csr<T, I, O>(m, n, nnz, rowptr, colind, values), has_aa, alpha_a,    \
         has_ax, alpha_x, x, n, y, m, inspect);                                    \
  }                                                                                \
  REF_API void ref_csc_spmv_##TN##_##IN##_##ON(                                    \
      int64_t m, int64_t n, int64_t nnz, const O* colptr, const I* rowind,         \
      const T* values, int has_aa, T alpha_a, int has_ax, T alpha_x, const T* x,   \
      T* y, int inspect) {                                                         \
    spmv(make_csc<T, I, O>(m, n, nnz, colptr, rowind, values), has_aa, alpha_a,    \
         has_ax, alpha_x, x, n, y, m, inspect);                                    \
  }                                                                                \
  REF_API void ref_csr_spmm_##TN##_##IN##_##ON(                                    \
      int64_t m, int64_t n, int64_t k, int64_t nnz, const O* rowptr,               \
      const I* colind, const T* values, int has_aa, T alpha_a, int has_ab,         \
      T alpha_b, const T* B, T* C, int inspect) {                                  \
    spmm<decltype(make_csr<T, I, O>(m, n, nnz, rowptr, colind, values)), T, I>(    \
        make_csr<T, I, O>(m, n, nnz, rowptr, colind, values), has_aa, alpha_a,     \
        has_ab, alpha_b, B, n, k, C, m, inspect);                                  \
  }                                                                                \
  REF_API void ref_csc_spmm_##TN##_##IN##_##ON(                                    \
      int64_t m, int64_t n, int64_t k, int64_t nnz, const O* colptr,               \
      const I* rowind, const T* values, int has_aa, T alpha_a, int has_ab,         \
      T alpha_b, const T* B, T* C, int inspect) {                                  \
    spmm<decltype(make_csc<T, I, O>(m, n, nnz, colptr, rowind, values)), T, I>(    \
        make_csc<T, I, O>(m, n, nnz, colptr, rowind, values), has_aa, alpha_a,     \
        has_ab, alpha_b, B, n, k, C, m, inspect);                                  \
  }

// transpose_inspect + transpose(info, a, b) exactly as test/gtest/transpose_test.cpp:33-34
#define DEF_TRANSPOSE(T, TN, I, IN, O, ON)                                         \
  REF_API void ref_csr_transpose_##TN##_##IN##_##ON(                               \
      int64_t m, int64_t n, int64_t nnz, const O* rowptr, const I* colind,         \
      const T* values, O* b_rowptr, I* b_colind, T* b_values) {                    \
    auto a = make_csr<T, I, O>(m, n, nnz, rowptr, colind, values);                 \
    spblas::csr_view<T, I, O> b(b_values, b_rowptr, b_colind,                      \
                                spblas::index<I>(I(n), I(m)), O(nnz));             \
    auto info = spblas::transpose_inspect(a, b);                                   \
    spblas::transpose(info, a, b);                                                 \
  }

// triangular_solve_inspect + triangular_solve as examples/sptrsv_csr.cpp:52-56 spells them
template <typename A, typename T>
void trsv(A a, int upper, int unit, int has_aa, T alpha_a, int has_ab, T alpha_b,
          const T* b, T* x, int64_t m) {
  std::span<T> bs(const_cast<T*>(b), size_t(m));
  std::span<T> xs(x, size_t(m));
  auto run2 = [&](auto&& av, auto&& bv) {
    auto go = [&](auto uplo, auto diag) {
      auto info = spblas::triangular_solve_inspect(av, uplo, diag, bv, xs);
      spblas::triangular_solve(info, av, uplo, diag, bv, xs);
    };
    if (upper && unit)
      go(spblas::upper_triangle_t{}, spblas::implicit_unit_diagonal_t{});
    else if (upper)
      go(spblas::upper_triangle_t{}, spblas::explicit_diagonal_t{});
    else if (unit)
      go(spblas::lower_triangle_t{}, spblas::implicit_unit_diagonal_t{});
    else
      go(spblas::lower_triangle_t{}, spblas::explicit_diagonal_t{});
  };
  if (has_aa && has_ab)
    run2(spblas::scaled(alpha_a, a), spblas::scaled(alpha_b, bs));
  else if (has_aa)
    run2(spblas::scaled(alpha_a, a), bs);
  else if (has_ab)
    run2(a, spblas::scaled(alpha_b, bs));
  else
    run2(a, bs);
}

#define DEF_TRSV(T, TN, I, IN, O, ON)                                              \
  REF_API void ref_csr_trsv_##TN##_##IN##_##ON(                                    \
      int64_t m, int64_t nnz, const O* rowptr, const I* colind, const T* values,   \
      int upper, int unit, int has_aa, T alpha_a, int has_ab, T alpha_b,           \
      const T* b, T* x) {                                                          \
    trsv(make_csr<T, I, O>(m, m, nnz, rowptr, colind, values), upper, unit,        \
         has_aa, alpha_a, has_ab, alpha_b, b, x, m);                               \
  }

DEF_TRSV(float, f32, int32_t, i32, int32_t, i32)
DEF_TRSV(float, f32, int32_t, i32, int64_t, i64)
DEF_TRSV(double, f64, int32_t, i32, int32_t, i32)
DEF_TRSV(double, f64, int32_t, i32, int64_t, i64)
DEF_TRSV(float, f32, int64_t, i64, int64_t, i64)

DEF_TRANSPOSE(float, f32, int32_t, i32, int32_t, i32)
DEF_TRANSPOSE(float, f32, int32_t, i32, int64_t, i64)
DEF_TRANSPOSE(double, f64, int32_t, i32, int32_t, i32)
DEF_TRANSPOSE(double, f64, int32_t, i32, int64_t, i64)
DEF_TRANSPOSE(int32_t, s32, int32_t, i32, int32_t, i32)
DEF_TRANSPOSE(float, f32, int64_t, i64, int64_t, i64)

DEF_OPS(float, f32, int32_t, i32, int32_t, i32)
DEF_OPS(float, f32, int32_t, i32, int64_t, i64)
DEF_OPS(double, f64, int32_t, i32, int32_t, i32)
DEF_OPS(double, f64, int32_t, i32, int64_t, i64)
DEF_OPS(int32_t, s32, int32_t, i32, int32_t, i32)
DEF_OPS(float, f32, int64_t, i64, int64_t, i64)

// ---- the reference's own fixtures (backend/generate.hpp:106-138,170-182) -------
#define DEF_GEN(T, TN)                                                             \
  REF_API void ref_generate_csr_##TN##_i32_i32(int64_t m, int64_t n, int64_t nnz,  \
                                               int64_t seed, T* values,            \
                                               int32_t* rowptr, int32_t* colind) { \
    auto [v, rp, ci, shape, nz] = spblas::generate_csr<T, int32_t, int32_t>(       \
        size_t(m), size_t(n), size_t(nnz), size_t(seed));                          \
    std::copy(v.begin(), v.end(), values);                                         \
    std::copy(rp.begin(), rp.end(), rowptr);                                       \
    std::copy(ci.begin(), ci.end(), colind);                                       \
  }                                                                                \
  REF_API void ref_generate_csc_##TN##_i32_i32(int64_t m, int64_t n, int64_t nnz,  \
                                               int64_t seed, T* values,            \
                                               int32_t* colptr, int32_t* rowind) { \
    auto [v, cp, ri, shape, nz] = spblas::generate_csc<T, int32_t, int32_t>(       \
        size_t(m), size_t(n), size_t(nnz), size_t(seed));                          \
    std::copy(v.begin(), v.end(), values);                                         \
    std::copy(cp.begin(), cp.end(), colptr);                                       \
    std::copy(ri.begin(), ri.end(), rowind);                                       \
  }                                                                                \
  REF_API void ref_generate_dense_##TN(int64_t m, int64_t n, int64_t seed,         \
                                       T* out) {                                   \
    auto [v, shape] = spblas::generate_dense<T>(size_t(m), size_t(n),              \
                                                size_t(seed));                     \
    std::copy(v.begin(), v.end(), out);                                            \
  }

DEF_GEN(float, f32)
DEF_GEN(double, f64)

REF_API int ref_abi_version(void) { return 1; }
