// compile_only.cpp — never run: every overload of the B200 backend instantiated with the operand
// spellings user code uses (lvalue, const lvalue and rvalue views; scaled / matrix_opt /
// transposed wrappers as temporaries; CSR and CSC; vectors and row-major matrices), so that a
// signature that only binds lvalues — as multiply_inspect(info, transposed(a), ...) did until the
// end of round 2 — fails HERE, in the CPU test suite (tests/test_cpp_compile.py, g++ -fsyntax-only),
// and not in a user's build.  The calls mirror the reference's own call sites:
// test/gtest/spmv_test.cpp, spmm_test.cpp, examples/simple_spmv.cpp, notes/spmv.hpp:12-22.
#include <span>
#include <utility>
#include <vector>

#include <spblas/spblas.hpp>

namespace {

template <typename T, typename I, typename O>
void spmv_forms(T* vals, O* ptr, I* ind, T* xp, T* yp, T* dp, I m, I n, O nnz) {
  using namespace spblas;
  csr_view<T, I, O> a(vals, ptr, ind, spblas::index<I>(m, n), nnz);
  const csr_view<T, I, O> ca = a;
  csc_view<T, I, O> c(vals, ptr, ind, spblas::index<I>(n, m), nnz);
  std::span<T> x(xp, n), y(yp, m), d(dp, m);
  std::span<const T> cx(xp, n);

  // -- no info
  multiply(a, x, y);
  multiply(ca, x, y);
  multiply(a, cx, y);
  multiply(scaled(2.0f, a), x, y);
  multiply(a, scaled(2.0f, x), y);
  multiply(scaled(2.0f, a), scaled(3.0f, x), y);
  multiply(matrix_opt(a), x, y);
  multiply(scaled(2.0f, matrix_opt(a)), x, y);
  multiply(transposed(c), x, y);          // csc(n x m) transposed -> csr(m x n)
  multiply(transposed(a), y, x);          // csr transposed -> csc(n x m)
  multiply(csr_view<T, I, O>(vals, ptr, ind, spblas::index<I>(m, n), nnz), x, y);
  multiply(a, x, y, d);
  multiply(scaled(2.0f, a), x, y, scaled(0.5f, d));
  multiply(a, x, y, scaled(2.0f, y));

  // -- inspect: every operand spelling, both overloads
  operation_info_t info = multiply_inspect(a, x, y);
  operation_info_t i2 = multiply_inspect(ca, x, y);
  operation_info_t i3 = multiply_inspect(scaled(2.0f, a), x, y);
  operation_info_t i4 = multiply_inspect(matrix_opt(a), x, y);
  operation_info_t i5 = multiply_inspect(transposed(a), y, x);
  operation_info_t i6 = multiply_inspect(matrix_opt(transposed(a)), y, x);
  operation_info_t i7 = multiply_inspect(a, scaled(2.0f, x), y);
  multiply_inspect(info, a, x, y);
  multiply_inspect(info, ca, x, y);
  multiply_inspect(info, scaled(2.0f, a), x, y);
  multiply_inspect(info, matrix_opt(a), x, y);
  multiply_inspect(info, transposed(a), y, x);
  multiply_inspect(info, matrix_opt(transposed(a)), y, x);
  matrix_opt a_opt(a);
  multiply_inspect(info, a_opt, x, y);
  auto at = transposed(a);
  multiply_inspect(info, at, y, x);

  // -- execute
  multiply(info, a, x, y);
  multiply(info, ca, cx, y);
  multiply(info, scaled(2.0f, a), x, y);
  multiply(info, a_opt, x, y);
  multiply(info, matrix_opt(a), x, y);
  multiply(info, transposed(a), y, x);
  multiply(info, at, y, x);
  multiply(info, a, x, y, d);
  multiply(info, scaled(2.0f, a), scaled(3.0f, x), y, scaled(-1.0f, d));
  multiply_execute(info, a, x, y);
  multiply_execute(info, scaled(2.0f, a), x, y);
  multiply_execute(info, transposed(a), y, x);
  multiply_execute(info, matrix_opt(transposed(a)), y, x);
  multiply_execute(info, a, x, y, d);
  multiply_execute(info, a, x, y, scaled(2.0f, d));
  operation_info_t moved = std::move(info);
  multiply_execute(moved, a, x, y);
  multiply_execute_host(moved, a, x, y, x, y);
  multiply_execute_host(moved, scaled(2.0f, a), x, y, x, y);
  (void)i2, (void)i3, (void)i4, (void)i5, (void)i6, (void)i7;
}

template <typename T, typename I, typename O>
void spmm_forms(T* vals, O* ptr, I* ind, T* bp, T* cp, T* dp, I m, I n, I k, O nnz) {
  using namespace spblas;
  csr_view<T, I, O> a(vals, ptr, ind, spblas::index<I>(m, n), nnz);
  const csr_view<T, I, O> ca = a;
  csc_view<T, I, O> cs(vals, ptr, ind, spblas::index<I>(m, n), nnz);
  mdspan_row_major<T, I> b(bp, n, k), c(cp, m, k), d(dp, m, k);

  multiply(a, b, c);
  multiply(ca, b, c);
  multiply(cs, b, c);
  multiply(scaled(2.0f, a), b, c);
  multiply(a, scaled(2.0f, b), c);
  multiply(matrix_opt(a), b, c);
  multiply(transposed(a), c, b);
  multiply(a, b, c, d);
  multiply(scaled(2.0f, a), b, c, scaled(0.5f, d));

  operation_info_t info = multiply_inspect(a, b, c);
  operation_info_t i2 = multiply_inspect(scaled(2.0f, a), b, c);
  operation_info_t i3 = multiply_inspect(matrix_opt(a), b, c);
  operation_info_t i4 = multiply_inspect(transposed(a), c, b);
  operation_info_t i5 = multiply_inspect(cs, b, c);
  multiply_inspect(info, a, b, c);
  multiply_inspect(info, ca, b, c);
  multiply_inspect(info, scaled(2.0f, a), b, c);
  multiply_inspect(info, matrix_opt(a), b, c);
  multiply_inspect(info, transposed(a), c, b);
  multiply_inspect(info, matrix_opt(cs), b, c);

  multiply(info, a, b, c);
  multiply(info, scaled(2.0f, a), b, c);
  multiply(info, matrix_opt(a), b, c);
  multiply(info, transposed(a), c, b);
  multiply(info, a, b, c, d);
  multiply(info, a, b, c, scaled(2.0f, d));
  multiply_execute(info, a, b, c);
  multiply_execute(info, scaled(2.0f, a), scaled(3.0f, b), c);
  multiply_execute(info, cs, b, c);
  multiply_execute(info, a, b, c, d);
  (void)i2, (void)i3, (void)i4, (void)i5;
}

template <typename T, typename I, typename O>
void other_forms(T* vals, O* ptr, I* ind, T* bv, O* bp, I* bi, T* xp, T* yp, I m, O nnz) {
  using namespace spblas;
  csr_view<T, I, O> a(vals, ptr, ind, spblas::index<I>(m, m), nnz);
  const csr_view<T, I, O> ca = a;
  csr_view<T, I, O> b(bv, bp, bi, spblas::index<I>(m, m), nnz);
  std::span<T> x(xp, m), y(yp, m);

  // transpose (algorithms/transpose_impl.hpp:9-60)
  operation_info_t ti = transpose_inspect(a, b);
  transpose(ti, a, b);
  transpose(a, b);
  transpose(ca, b);
  transpose(ti, ca, csr_view<T, I, O>(bv, bp, bi, spblas::index<I>(m, m), nnz));

  // triangular_solve (algorithms/triangular_solve_impl.hpp:14-107)
  operation_info_t si = triangular_solve_inspect(a, lower_triangle, explicit_diagonal, y, x);
  triangular_solve_inspect(si, a, upper_triangle, implicit_unit_diagonal, y, x);
  triangular_solve(si, a, lower_triangle, explicit_diagonal, y, x);
  triangular_solve(si, ca, lower_triangle, explicit_diagonal, scaled(2.0f, y), x);
  triangular_solve(si, scaled(2.0f, a), upper_triangle, implicit_unit_diagonal, y, x);
  triangular_solve(a, lower_triangle, explicit_diagonal, y, x);
  triangular_solve(matrix_opt(a), lower_triangle, implicit_unit_diagonal, scaled(2.0f, y), x);
  matrix_opt a_opt(a);
  triangular_solve(a_opt, upper_triangle, explicit_diagonal, y, x);
}

} // namespace

template void other_forms<float, spblas::index_t, spblas::offset_t>(
    float*, spblas::offset_t*, spblas::index_t*, float*, spblas::offset_t*, spblas::index_t*, float*,
    float*, spblas::index_t, spblas::offset_t);
template void other_forms<double, std::int32_t, std::int64_t>(double*, std::int64_t*, std::int32_t*,
                                                             double*, std::int64_t*, std::int32_t*,
                                                             double*, double*, std::int32_t,
                                                             std::int64_t);

// explicit instantiations: the reference's device-test types, 64-bit offsets, 64-bit everything
template void spmv_forms<float, spblas::index_t, spblas::offset_t>(
    float*, spblas::offset_t*, spblas::index_t*, float*, float*, float*, spblas::index_t,
    spblas::index_t, spblas::offset_t);
template void spmv_forms<double, std::int32_t, std::int64_t>(double*, std::int64_t*, std::int32_t*,
                                                            double*, double*, double*, std::int32_t,
                                                            std::int32_t, std::int64_t);
template void spmv_forms<float, std::int64_t, std::int64_t>(float*, std::int64_t*, std::int64_t*,
                                                           float*, float*, float*, std::int64_t,
                                                           std::int64_t, std::int64_t);
template void spmm_forms<float, spblas::index_t, spblas::offset_t>(
    float*, spblas::offset_t*, spblas::index_t*, float*, float*, float*, spblas::index_t,
    spblas::index_t, spblas::index_t, spblas::offset_t);
template void spmm_forms<double, std::int32_t, std::int64_t>(double*, std::int64_t*, std::int32_t*,
                                                            double*, double*, double*, std::int32_t,
                                                            std::int32_t, std::int32_t, std::int64_t);

int main() {
  return 0;
}
