// dropin_test.cpp — the B200 backend used through the UNMODIFIED reference API:
//   #include <spblas/spblas.hpp> compiled with -DSPBLAS_ENABLE_B200 against the reference's
//   own headers (+ integration/enable_b200.patch) and include/spblas/vendor/b200/.
// gtest is not available offline, so this is a small self-contained harness.  The cases
// re-host what the reference's GPU contract test checks
// (test/gtest/device/spmv_test.cpp:11-146: SpMV, SpMV_Ascaled, SpMV_BScaled on util::dims with
// generate_csr seed 0, judged by EXPECT_EQ_ of test/gtest/util.hpp:7-23) and add what the
// reference has no device test for: SpMM (cf. test/gtest/spmm_test.cpp:6-221), CSC,
// transposed(), matrix_opt, multiply_inspect + multiply(info, ..) / multiply_execute, fp64,
// 64-bit offsets, and the error behaviour.
#include <spblas/spblas.hpp>

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <span>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace {

int g_checks = 0, g_failures = 0;
std::string g_case;

void fail(const std::string& what) {
  ++g_failures;
  std::printf("  FAIL [%s] %s\n", g_case.c_str(), what.c_str());
}

#define CUDA_OK(expr)                                                              \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__,  \
                  __LINE__);                                                       \
      std::exit(2);                                                                \
    }                                                                              \
  } while (0)

// the reference tests' acceptance criterion (test/gtest/util.hpp:7-23)
template <typename T>
bool close_enough(T t, T u) {
  if constexpr (std::is_floating_point_v<T>) {
    const T eps = 64 * std::numeric_limits<T>::epsilon();
    const T norm = std::min(std::abs(t) + std::abs(u), std::numeric_limits<T>::max());
    return std::abs(t - u) <= std::max(std::numeric_limits<T>::min(), eps * norm);
  } else {
    return t == u;
  }
}

template <typename T>
void expect_all_close(const std::vector<T>& ref, const std::vector<T>& got) {
  ++g_checks;
  if (ref.size() != got.size()) {
    fail("size mismatch");
    return;
  }
  for (std::size_t i = 0; i < ref.size(); ++i) {
    if (!close_enough(ref[i], got[i])) {
      fail("element " + std::to_string(i) + ": expected " + std::to_string(ref[i]) +
           ", got " + std::to_string(got[i]));
      return;
    }
  }
}

template <typename T>
class device_array {
public:
  device_array() = default;
  explicit device_array(const std::vector<T>& h) : n_(h.size()) {
    CUDA_OK(cudaMalloc(&p_, std::max<std::size_t>(n_, 1) * sizeof(T)));
    if (n_)
      CUDA_OK(cudaMemcpy(p_, h.data(), n_ * sizeof(T), cudaMemcpyHostToDevice));
  }
  device_array(std::size_t n, T fill) : device_array(std::vector<T>(n, fill)) {}
  device_array(const device_array&) = delete;
  device_array& operator=(const device_array&) = delete;
  ~device_array() {
    if (p_)
      cudaFree(p_);
  }
  T* get() const { return p_; }
  std::size_t size() const { return n_; }
  std::vector<T> to_host() const {
    std::vector<T> h(n_);
    if (n_)
      CUDA_OK(cudaMemcpy(h.data(), p_, n_ * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }

private:
  T* p_ = nullptr;
  std::size_t n_ = 0;
};

const std::vector<std::tuple<int, int, int>> dims = {
    {1000, 100, 100}, {100, 1000, 10000}, {40, 40, 1000}}; // util::dims

template <typename T, typename I, typename O>
std::vector<T> host_spmv(int m, const std::vector<O>& rowptr, const std::vector<I>& colind,
                         const std::vector<T>& values, const std::vector<T>& x, T alpha) {
  std::vector<T> y(m, 0);
  for (int i = 0; i < m; ++i)
    for (O p = rowptr[i]; p < rowptr[i + 1]; ++p)
      y[i] += alpha * values[p] * x[colind[p]];
  return y;
}

template <typename T, typename I, typename O>
std::vector<T> host_spmm(int m, int k, const std::vector<O>& rowptr,
                         const std::vector<I>& colind, const std::vector<T>& values,
                         const std::vector<T>& B, T alpha) {
  std::vector<T> C(std::size_t(m) * k, 0);
  for (int i = 0; i < m; ++i)
    for (O p = rowptr[i]; p < rowptr[i + 1]; ++p)
      for (int j = 0; j < k; ++j)
        C[std::size_t(i) * k + j] += alpha * values[p] * B[std::size_t(colind[p]) * k + j];
  return C;
}

// ---- SpMV: the three reference device tests + inspect/execute reuse ------------------
template <typename T, typename I, typename O>
void spmv_cases() {
  for (auto [m, n, nnz] : dims) {
    auto [values, rowptr, colind, shape, nnz_] = spblas::generate_csr<T, I, O>(m, n, nnz);
    std::vector<T> x(n, 1);
    device_array<T> d_values(values);
    device_array<O> d_rowptr(rowptr);
    device_array<I> d_colind(colind);
    device_array<T> d_x(x);
    device_array<T> d_y(m, std::numeric_limits<T>::quiet_NaN()); // stale y must vanish
    spblas::csr_view<T, I, O> a(d_values.get(), d_rowptr.get(), d_colind.get(), shape,
                                O(nnz));
    std::span<T> x_span(d_x.get(), n);
    std::span<T> y_span(d_y.get(), m);

    g_case = "SpMV";
    spblas::multiply(a, x_span, y_span);
    expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(1)), d_y.to_host());

    for (int alpha : {-10, 1, 5}) {
      g_case = "SpMV_Ascaled";
      spblas::multiply(spblas::scaled(alpha, a), x_span, y_span);
      expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(alpha)),
                       d_y.to_host());
      g_case = "SpMV_BScaled";
      spblas::multiply(a, spblas::scaled(alpha, x_span), y_span);
      expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(alpha)),
                       d_y.to_host());
    }

    g_case = "SpMV inspect/execute";
    auto info = spblas::multiply_inspect(a, x_span, y_span);
    for (int rep = 0; rep < 3; ++rep) {
      spblas::multiply(info, spblas::scaled(2.0f, a), x_span, y_span);
      expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(2)),
                       d_y.to_host());
    }
    spblas::multiply_execute(info, a, x_span, y_span);
    expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(1)), d_y.to_host());
    spblas::operation_info_t moved = std::move(info); // move-only state travels
    spblas::multiply(moved, a, x_span, y_span);
    expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(1)), d_y.to_host());

    g_case = "SpMV host vectors (pipelined upload / kernels / download)";
    {
      std::vector<T> y_host(m, T(77));
      device_array<T> x_stage(std::size_t(n), T(0)), y_stage(std::size_t(m), T(0));
      std::span<T> xs(x_stage.get(), n), ys(y_stage.get(), m);
      std::span<T> xh(x.data(), n), yh(y_host.data(), m);
      spblas::multiply_execute_host(moved, spblas::scaled(2.0f, a), xh, yh, xs, ys);
      CUDA_OK(cudaDeviceSynchronize());
      spblas::multiply(moved, spblas::scaled(2.0f, a), x_span, y_span);
      ++g_checks;
      if (y_host != d_y.to_host())
        fail("host-vector execute differs from the device-vector execute");
      expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(2)), y_host);
    }

    g_case = "SpMV 4-argument form: y = alpha A x + beta d";
    {
      std::vector<T> d(m);
      for (int i = 0; i < m; ++i)
        d[i] = T(1 + i % 5);
      device_array<T> d_d(d);
      std::span<T> d_span(d_d.get(), m);
      auto expect = [&](T alpha, T beta, const std::vector<T>& dd) {
        std::vector<T> r = host_spmv<T, I, O>(m, rowptr, colind, values, x, alpha);
        for (int i = 0; i < m; ++i)
          r[i] += beta * dd[i];
        return r;
      };
      spblas::multiply(a, x_span, y_span, d_span);                        // beta = 1, no info
      expect_all_close(expect(T(1), T(1), d), d_y.to_host());
      spblas::multiply(moved, spblas::scaled(2.0f, a), x_span, y_span,
                       spblas::scaled(-3.0f, d_span));                     // alpha = 2, beta = -3
      expect_all_close(expect(T(2), T(-3), d), d_y.to_host());
      spblas::multiply_execute(moved, a, x_span, y_span, spblas::scaled(0.5f, d_span));
      expect_all_close(expect(T(1), T(0.5), d), d_y.to_host());
      // in place: y = A x + 2 y (the solver update)
      CUDA_OK(cudaMemcpy(d_y.get(), d.data(), m * sizeof(T), cudaMemcpyHostToDevice));
      spblas::multiply(moved, a, x_span, y_span, spblas::scaled(2.0f, y_span));
      expect_all_close(expect(T(1), T(2), d), d_y.to_host());
      // a second no-info call on the same structure: the cached plan, checked on the device
      spblas::multiply(a, x_span, y_span, spblas::scaled(2.0f, d_span));
      expect_all_close(expect(T(1), T(2), d), d_y.to_host());
    }

    g_case = "SpMV matrix_opt";
    spblas::matrix_opt a_opt(a);
    auto info2 = spblas::multiply_inspect(a_opt, x_span, y_span);
    spblas::multiply(info2, a_opt, x_span, y_span);
    expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(1)), d_y.to_host());
  }
}

// ---- CSC and transposed() -----------------------------------------------------------------
void csc_cases() {
  using T = float;
  using I = spblas::index_t;
  using O = spblas::offset_t;
  for (auto [m, n, nnz] : dims) {
    auto [values, colptr, rowind, shape, nnz_] = spblas::generate_csc<T, I, O>(m, n, nnz);
    std::vector<T> x(n);
    for (int j = 0; j < n; ++j)
      x[j] = T(1 + j % 3);
    std::vector<T> ref(m, 0);
    for (int j = 0; j < n; ++j)
      for (O p = colptr[j]; p < colptr[j + 1]; ++p)
        ref[rowind[p]] += values[p] * x[j];
    device_array<T> d_values(values);
    device_array<O> d_colptr(colptr);
    device_array<I> d_rowind(rowind);
    device_array<T> d_x(x), d_y(m, T(-1));
    spblas::csc_view<T, I, O> a(d_values.get(), d_colptr.get(), d_rowind.get(), shape, O(nnz));
    std::span<T> x_span(d_x.get(), n), y_span(d_y.get(), m);
    g_case = "CscView SpMV";
    spblas::multiply(a, x_span, y_span);
    expect_all_close(ref, d_y.to_host());
    auto info = spblas::multiply_inspect(a, x_span, y_span);
    spblas::multiply(info, a, x_span, y_span);
    expect_all_close(ref, d_y.to_host());

    // transposed(csr) is the csc over the same arrays (algorithms/transposed.hpp:7-13):
    // read the colptr/rowind arrays as the CSR of A^T (n x m) and transpose that view
    // back, which must give A again
    g_case = "transposed";
    spblas::csr_view<T, I, O> at_csr(d_values.get(), d_colptr.get(), d_rowind.get(),
                                     spblas::index<I>(I(n), I(m)), O(nnz));
    auto a_again = spblas::transposed(at_csr);
    device_array<T> d_y2(m, T(0));
    spblas::multiply(a_again, x_span, std::span<T>(d_y2.get(), m));
    expect_all_close(ref, d_y2.to_host());
    // and the CSR view itself computes A^T x2
    std::vector<T> x2(m, 1), ref2(n, 0);
    for (int j = 0; j < n; ++j)
      for (O p = colptr[j]; p < colptr[j + 1]; ++p)
        ref2[j] += values[p] * x2[rowind[p]];
    device_array<T> d_x2(x2), d_y3(n, T(0));
    spblas::multiply(at_csr, std::span<T>(d_x2.get(), m), std::span<T>(d_y3.get(), n));
    expect_all_close(ref2, d_y3.to_host());
  }
}

// ---- one info, two operands: the alternating loop of the reference's notes ---------------------
// notes/spmv.hpp:12-22: multiply_inspect(info, a, x, y); multiply_inspect(info, transposed(a),
// y, x); then multiply_execute(info, a, x, y) and multiply_execute(info, transposed(a), y, x) in
// turn.  The info keeps one plan per inspected structure: the executes switch plans, they do
// not re-inspect (the plan pointers seen by the two operands never change).
void alternating_case() {
  using T = double;
  using I = spblas::index_t;
  using O = spblas::offset_t;
  const int m = 300, n = 200, nnz = 4000;
  auto [values, rowptr, colind, shape, nnz_] = spblas::generate_csr<T, I, O>(m, n, nnz);
  for (auto& v : values)
    v = T(0.01) * v; // keeps the iterates small
  device_array<T> d_values(values);
  device_array<O> d_rowptr(rowptr);
  device_array<I> d_colind(colind);
  std::vector<T> x(n);
  for (int j = 0; j < n; ++j)
    x[j] = T(1 + j % 5);
  device_array<T> d_x(x), d_y(m, T(-1));
  spblas::csr_view<T, I, O> a(d_values.get(), d_rowptr.get(), d_colind.get(), shape, O(nnz));
  std::span<T> xs(d_x.get(), n), ys(d_y.get(), m);
  g_case = "alternating A x / A^T y with one info";
  spblas::operation_info_t info;
  spblas::multiply_inspect(info, a, xs, ys);
  spblas_b200_plan* plan_a = info.state_.plan();
  spblas::multiply_inspect(info, spblas::transposed(a), ys, xs);
  spblas_b200_plan* plan_at = info.state_.plan();
  ++g_checks;
  if (plan_a == nullptr || plan_at == nullptr || plan_a == plan_at)
    fail("the two operands must own two plans");
  for (int it = 0; it < 3; ++it) {
    spblas::multiply_execute(info, a, xs, ys);
    ++g_checks;
    if (info.state_.plan() != plan_a)
      fail("execute on a re-inspected or replaced plan (A)");
    std::vector<T> y_ref = host_spmv<T, I, O>(m, rowptr, colind, values, x, T(1));
    expect_all_close(y_ref, d_y.to_host());
    spblas::multiply_execute(info, spblas::transposed(a), ys, xs);
    ++g_checks;
    if (info.state_.plan() != plan_at)
      fail("execute on a re-inspected or replaced plan (A^T)");
    std::vector<T> x_ref(n, T(0));
    for (int i = 0; i < m; ++i)
      for (O p = rowptr[i]; p < rowptr[i + 1]; ++p)
        x_ref[colind[p]] += values[p] * y_ref[i];
    expect_all_close(x_ref, d_x.to_host());
    x = x_ref;
  }
  // a moved info takes every plan with it
  spblas::operation_info_t moved = std::move(info);
  spblas::multiply_execute(moved, a, xs, ys);
  ++g_checks;
  if (moved.state_.plan() != plan_a || info.state_.plan() != nullptr)
    fail("move must carry the parked plans");
  expect_all_close(host_spmv<T, I, O>(m, rowptr, colind, values, x, T(1)), d_y.to_host());
}

// ---- SpMM ---------------------------------------------------------------------------------------
template <typename T>
void spmm_cases() {
  using I = spblas::index_t;
  for (auto [m, k, nnz] : dims) {
    for (int n : {1, 8, 32, 64, 512}) {
      auto [values, rowptr, colind, shape, nnz_] = spblas::generate_csr<T, I>(m, k, nnz);
      auto [b_values, b_shape] = spblas::generate_dense<T>(k, n);
      device_array<T> d_values(values);
      device_array<I> d_rowptr(rowptr), d_colind(colind);
      device_array<T> d_b(b_values), d_c(std::size_t(m) * n, std::numeric_limits<T>::quiet_NaN());
      spblas::csr_view<T, I> a(d_values.get(), d_rowptr.get(), d_colind.get(), shape, I(nnz));
      spblas::mdspan_row_major<T, I> b(d_b.get(), k, n);
      spblas::mdspan_row_major<T, I> c(d_c.get(), m, n);

      g_case = "SpMM n=" + std::to_string(n);
      spblas::multiply(a, b, c);
      expect_all_close(host_spmm<T, I, I>(m, n, rowptr, colind, values, b_values, T(1)),
                       d_c.to_host());
      g_case = "SpMM_AScaled n=" + std::to_string(n);
      spblas::multiply(spblas::scaled(2.0f, a), b, c);
      expect_all_close(host_spmm<T, I, I>(m, n, rowptr, colind, values, b_values, T(2)),
                       d_c.to_host());
      g_case = "SpMM_BScaled n=" + std::to_string(n);
      spblas::multiply(a, spblas::scaled(2.0f, b), c);
      expect_all_close(host_spmm<T, I, I>(m, n, rowptr, colind, values, b_values, T(2)),
                       d_c.to_host());
      g_case = "SpMM_Aopt n=" + std::to_string(n);
      spblas::matrix_opt a_opt(a);
      auto info = spblas::multiply_inspect(a_opt, b, c); // examples/spmm_csr.cpp:45-46
      spblas::multiply(info, a_opt, b, c);
      expect_all_close(host_spmm<T, I, I>(m, n, rowptr, colind, values, b_values, T(1)),
                       d_c.to_host());
      spblas::multiply_execute(info, a_opt, b, c);
      expect_all_close(host_spmm<T, I, I>(m, n, rowptr, colind, values, b_values, T(1)),
                       d_c.to_host());
      g_case = "SpMM 4-argument form n=" + std::to_string(n);
      {
        std::vector<T> dd(std::size_t(m) * n);
        for (std::size_t i = 0; i < dd.size(); ++i)
          dd[i] = T(1 + i % 7);
        device_array<T> d_d(dd);
        spblas::mdspan_row_major<T, I> d(d_d.get(), m, n);
        auto expect = [&](T alpha, T beta) {
          std::vector<T> r = host_spmm<T, I, I>(m, n, rowptr, colind, values, b_values, alpha);
          for (std::size_t i = 0; i < r.size(); ++i)
            r[i] += beta * dd[i];
          return r;
        };
        spblas::multiply(a, b, c, d);
        expect_all_close(expect(T(1), T(1)), d_c.to_host());
        spblas::multiply(info, spblas::scaled(2.0f, a_opt), b, c, spblas::scaled(-1.5f, d));
        expect_all_close(expect(T(2), T(-1.5)), d_c.to_host());
        CUDA_OK(cudaMemcpy(d_c.get(), dd.data(), dd.size() * sizeof(T), cudaMemcpyHostToDevice));
        spblas::multiply_execute(info, a_opt, b, c, spblas::scaled(3.0f, c)); // in place
        expect_all_close(expect(T(1), T(3)), d_c.to_host());
      }
    }
  }
}

void csc_spmm_case() {
  using T = float;
  using I = spblas::index_t;
  auto [m, k, nnz] = dims[1];
  const int n = 32;
  auto [values, colptr, rowind, shape, nnz_] = spblas::generate_csc<T, I>(m, k, nnz);
  auto [b_values, b_shape] = spblas::generate_dense<T>(k, n);
  std::vector<T> ref(std::size_t(m) * n, 0);
  for (int j = 0; j < k; ++j)
    for (I p = colptr[j]; p < colptr[j + 1]; ++p)
      for (int c = 0; c < n; ++c)
        ref[std::size_t(rowind[p]) * n + c] += values[p] * b_values[std::size_t(j) * n + c];
  device_array<T> d_values(values), d_b(b_values), d_c(std::size_t(m) * n, T(0));
  device_array<I> d_colptr(colptr), d_rowind(rowind);
  spblas::csc_view<T, I> a(d_values.get(), d_colptr.get(), d_rowind.get(), shape, I(nnz));
  g_case = "CscView SpMM";
  spblas::multiply(a, spblas::mdspan_row_major<T, I>(d_b.get(), k, n),
                   spblas::mdspan_row_major<T, I>(d_c.get(), m, n));
  expect_all_close(ref, d_c.to_host());
}

// ---- transpose(a, b): the body of test/gtest/transpose_test.cpp on device memory --------------
template <typename T, typename I, typename O>
void transpose_cases() {
  for (auto [m, n, nnz] : dims) {
    auto [values, rowptr, colind, shape, nnz_] = spblas::generate_csr<T, I, O>(m, n, nnz);
    device_array<T> d_values(values);
    device_array<O> d_rowptr(rowptr);
    device_array<I> d_colind(colind);
    spblas::csr_view<T, I, O> a(d_values.get(), d_rowptr.get(), d_colind.get(), shape, O(nnz));
    device_array<T> d_bv(std::size_t(nnz), T(0));
    device_array<O> d_brp(std::size_t(n) + 1, O(-1));
    device_array<I> d_bci(std::size_t(nnz), I(-1));
    spblas::csr_view<T, I, O> b(d_bv.get(), d_brp.get(), d_bci.get(),
                                spblas::index<I>(I(n), I(m)), O(nnz));
    g_case = "transpose";
    auto info = spblas::transpose_inspect(a, b);
    spblas::transpose(info, a, b);
    // the reference's host algorithm (algorithms/transpose_impl.hpp:33-50), restated
    std::vector<O> rp(n + 1, 0);
    std::vector<I> ci(nnz);
    std::vector<T> v(nnz);
    for (int i = 0; i < m; ++i)
      for (O p = rowptr[i]; p < rowptr[i + 1]; ++p)
        rp[colind[p] + 1]++;
    O run = 0;
    for (int j = 0; j <= n; ++j) {
      const O c = rp[j];
      rp[j] = run;
      run += c;
    }
    for (int i = 0; i < m; ++i)
      for (O p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        const O out = rp[colind[p] + 1]++;
        ci[out] = I(i);
        v[out] = values[p];
      }
    ++g_checks;
    if (d_brp.to_host() != rp || d_bci.to_host() != ci || d_bv.to_host() != v)
      fail("transpose differs from the reference algorithm");
    ++g_checks;
    if (std::size_t(b.size()) != std::size_t(nnz))
      fail("b.update() did not set the stored-entry count");
    // the overload without info, and B used as an operand afterwards: B x = A^T x
    spblas::transpose(a, b);
    std::vector<T> x(m, T(1));
    device_array<T> d_x(x), d_y(std::size_t(n), T(0)), d_y2(std::size_t(n), T(0));
    spblas::multiply(b, std::span<T>(d_x.get(), m), std::span<T>(d_y.get(), n));
    spblas::multiply(spblas::transposed(a), std::span<T>(d_x.get(), m),
                     std::span<T>(d_y2.get(), n));
    expect_all_close(d_y.to_host(), d_y2.to_host());
  }
  g_case = "transpose errors";
  auto [values, rowptr, colind, shape, nnz] = spblas::generate_csr<T, I, O>(40, 30, 100);
  device_array<T> d_values(values), d_bv(std::size_t(99), T(0));
  device_array<O> d_rowptr(rowptr), d_brp(std::size_t(31), O(0));
  device_array<I> d_colind(colind), d_bci(std::size_t(99), I(0));
  spblas::csr_view<T, I, O> a(d_values.get(), d_rowptr.get(), d_colind.get(), shape, O(100));
  ++g_checks;
  try {
    spblas::csr_view<T, I, O> bad(d_bv.get(), d_brp.get(), d_bci.get(),
                                  spblas::index<I>(I(31), I(40)), O(99));
    spblas::transpose(a, bad);
    fail("no exception for incompatible dimensions");
  } catch (const std::invalid_argument& e) {
    if (std::string(e.what()) != "transpose: matrix dimensions are incompatible.")
      fail(std::string("unexpected message: ") + e.what());
  }
  ++g_checks;
  try {
    spblas::csr_view<T, I, O> small(d_bv.get(), d_brp.get(), d_bci.get(),
                                    spblas::index<I>(I(30), I(40)), O(99));
    spblas::transpose(a, small);
    fail("no exception for an output that is too small");
  } catch (const std::runtime_error& e) {
    if (std::string(e.what()) != "transpose: Transpose ran out of memory.")
      fail(std::string("unexpected message: ") + e.what());
  }
}

// ---- triangular_solve: the reference's host loop restated, results must be IDENTICAL -----------
template <typename T, typename I, typename O, typename Triangle, typename Diag>
void trsv_case(Triangle uplo, Diag diag) {
  constexpr bool upper = std::is_same_v<Triangle, spblas::upper_triangle_t>;
  constexpr bool unit = std::is_same_v<Diag, spblas::implicit_unit_diagonal_t>;
  for (auto [m, n, nnz] : std::vector<std::tuple<int, int, int>>{
           {1000, 1000, 100}, {100, 100, 100}, {40, 40, 1000}}) { // util::square_dims
    auto [gv, grp, gci, shape, nnz_] = spblas::generate_csr<T, I, O>(m, n, nnz);
    // put a diagonal entry in front of every row; off-diagonals scaled as the reference test does
    std::vector<T> values;
    std::vector<O> rowptr(m + 1, 0);
    std::vector<I> colind;
    for (int i = 0; i < m; ++i) {
      colind.push_back(I(i));
      values.push_back(T(2 + i % 3));
      for (O p = grp[i]; p < grp[i + 1]; ++p) {
        colind.push_back(gci[p]);
        values.push_back(T(1e-3) * gv[p]);
      }
      rowptr[i + 1] = O(colind.size());
    }
    std::vector<T> b(m), x_ref(m, T(0));
    for (int i = 0; i < m; ++i)
      b[i] = T(1 + i % 7);
    const T alpha = T(1.2);
    T dval = 0;
    for (int s = 0; s < m; ++s) {                 // algorithms/triangular_solve_impl.hpp:61-93
      const int i = upper ? m - 1 - s : s;
      T dot = 0;
      for (O p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        const int k = int(colind[p]);
        if (upper ? k > i : k < i) {
          volatile T prod = values[p] * x_ref[k]; // no contraction: multiply, then add
          dot = dot + prod;
        } else if (k == i) {
          dval = values[p];
        }
      }
      volatile T bi = alpha * b[i];
      volatile T num = bi - dot;
      x_ref[i] = unit ? T(num) : T(num / dval);
    }
    device_array<T> d_values(values), d_b(b), d_x(std::size_t(m), std::numeric_limits<T>::quiet_NaN());
    device_array<O> d_rowptr(rowptr);
    device_array<I> d_colind(colind);
    spblas::csr_view<T, I, O> a(d_values.get(), d_rowptr.get(), d_colind.get(),
                                spblas::index<I>(I(m), I(m)), O(colind.size()));
    std::span<T> b_span(d_b.get(), m), x_span(d_x.get(), m);
    g_case = "triangular_solve";
    auto info = spblas::triangular_solve_inspect(a, uplo, diag, spblas::scaled(alpha, b_span), x_span);
    spblas::triangular_solve(info, a, uplo, diag, spblas::scaled(alpha, b_span), x_span);
    ++g_checks;
    if (d_x.to_host() != x_ref)
      fail("triangular_solve differs from the reference loop");
    spblas::matrix_opt a_opt(a);
    spblas::triangular_solve(a_opt, uplo, diag, spblas::scaled(alpha, b_span), x_span); // no info
    ++g_checks;
    if (d_x.to_host() != x_ref)
      fail("triangular_solve (no info, matrix_opt) differs from the reference loop");
  }
}

void trsv_errors() {
  using T = float;
  using I = spblas::index_t;
  g_case = "triangular_solve errors";
  std::vector<T> v{1, 1, 1};
  std::vector<I> rp{0, 1, 2, 3}, ci{0, 0, 2};
  device_array<T> d_v(v), d_b(std::size_t(3), T(1)), d_x(std::size_t(3), T(0));
  device_array<I> d_rp(rp), d_ci(ci);
  spblas::csr_view<T, I> a(d_v.get(), d_rp.get(), d_ci.get(), spblas::index<I>(3, 3), 3);
  ++g_checks;
  try {
    spblas::triangular_solve(a, spblas::lower_triangle, spblas::explicit_diagonal,
                             std::span<T>(d_b.get(), 3), std::span<T>(d_x.get(), 3));
    fail("no exception for a row without a diagonal under explicit_diagonal");
  } catch (const std::runtime_error&) {
  }
  spblas::triangular_solve(a, spblas::lower_triangle, spblas::implicit_unit_diagonal,
                           std::span<T>(d_b.get(), 3), std::span<T>(d_x.get(), 3));
  expect_all_close(std::vector<T>{1, 0, 1}, d_x.to_host());
}

// ---- error behaviour --------------------------------------------------------------------------
void error_cases() {
  using T = float;
  using I = spblas::index_t;
  auto [values, rowptr, colind, shape, nnz] = spblas::generate_csr<T, I>(40, 40, 100);
  device_array<T> d_values(values), d_x(41, T(1)), d_y(40, T(0));
  device_array<I> d_rowptr(rowptr), d_colind(colind);
  spblas::csr_view<T, I> a(d_values.get(), d_rowptr.get(), d_colind.get(), shape, nnz);
  g_case = "shape mismatch";
  ++g_checks;
  try {
    spblas::multiply(a, std::span<T>(d_x.get(), 41), std::span<T>(d_y.get(), 40));
    fail("no exception for a vector of the wrong length");
  } catch (const std::invalid_argument& e) {
    if (std::string(e.what()) != "multiply: matrix and vector dimensions are incompatible.")
      fail(std::string("unexpected message: ") + e.what());
  }
  // conjugated() of a real matrix is the matrix itself (algorithms/conjugated_impl.hpp:21-28),
  // so it must simply compute; complex scalars have no B200 overload at all (type gate).
  g_case = "conjugated(real) is the identity";
  {
    std::vector<T> ones(40, T(1));
    device_array<T> d_ones(ones);
    spblas::multiply(spblas::conjugated(a), std::span<T>(d_ones.get(), 40),
                     std::span<T>(d_y.get(), 40));
    expect_all_close(host_spmv<T, I, I>(40, rowptr, colind, values, ones, T(1)),
                     d_y.to_host());
  }
}

} // namespace

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    std::printf("dropin_test: no CUDA device\n");
    return 77;
  }
  static_assert(std::is_same_v<spblas::index_t, std::int32_t>,
                "the B200 backend sets 32-bit default indices like the other GPU backends");
  spmv_cases<float, spblas::index_t, spblas::offset_t>(); // the reference's device test types
  spmv_cases<double, std::int32_t, std::int64_t>();
  spmv_cases<float, std::int64_t, std::int64_t>();
  csc_cases();
  alternating_case();
  spmm_cases<float>();
  spmm_cases<double>();
  csc_spmm_case();
  transpose_cases<float, spblas::index_t, spblas::offset_t>();
  transpose_cases<double, std::int32_t, std::int64_t>();
  trsv_case<float, spblas::index_t, spblas::offset_t>(spblas::lower_triangle, spblas::implicit_unit_diagonal);
  trsv_case<float, spblas::index_t, spblas::offset_t>(spblas::upper_triangle, spblas::explicit_diagonal);
  trsv_case<double, std::int32_t, std::int64_t>(spblas::lower_triangle, spblas::explicit_diagonal);
  trsv_case<double, std::int32_t, std::int64_t>(spblas::upper_triangle, spblas::implicit_unit_diagonal);
  trsv_errors();
  error_cases();
  std::printf("dropin_test: %d checks, %d failures\n", g_checks, g_failures);
  return g_failures == 0 ? 0 : 1;
}
