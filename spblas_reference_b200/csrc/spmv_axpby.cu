// spmv_axpby.cu — the 4-argument product y = alpha * A * x + beta * d (spblas_b200_spmv_axpby;
// reference convention: vendor/rocsparse/multiply_spgemm.hpp:69-118): the ADD = true
// instantiations of the kernels in spmv_kernels.cuh, compiled apart from the plain product's.
#include "spmv_kernels.cuh"

namespace b200 {

int spmv_dispatch_addend(spblas_b200_plan* p, int val_type, int variant, const void* alpha,
                         const void* values, const void* x, void* y, int64_t T0, int64_t T1) {
  switch (val_type) {
  case SPBLAS_B200_F32:
    return dispatch_index<float, true>(p, variant, alpha, values, x, y, T0, T1);
  case SPBLAS_B200_F64:
    return dispatch_index<double, true>(p, variant, alpha, values, x, y, T0, T1);
  case SPBLAS_B200_S32:
    return dispatch_index<int32_t, true>(p, variant, alpha, values, x, y, T0, T1);
  default:
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "unknown value type");
  }
}

} // namespace b200
