// spmv.cu — y = alpha * A * x for the effective CSR structure of a plan: which kernel runs
// (prepare_spmv), and the 3-argument product's instantiations of the kernels in
// spmv_kernels.cuh.  The 4-argument product's live in spmv_axpby.cu.
#include "spmv_kernels.cuh"

namespace b200 {

// spmv_axpby.cu
int spmv_dispatch_addend(spblas_b200_plan* p, int val_type, int variant, const void* alpha,
                         const void* values, const void* x, void* y, int64_t T0, int64_t T1);


// Which kernel runs this product, and the partition it runs on.  The pipelined kernel
// wins when most tiles are uniform (stencils, fixed-degree graphs: coalesced row-ordered
// gathers, no row-end lookups); matrices with mixed row lengths are bound by random
// gathers of x, where autonomous warps keep the L1 tag stage busiest (warp streams).
// Arrays that are not 16-byte aligned fall back to the one-tile-per-CTA kernel.
int prepare_spmv(spblas_b200_plan* p, int val_type, const void* values, int* variant,
                 const int64_t** starts, int64_t* units) {
  const bool forced = p->forced_variant >= 0;
  int v = forced ? p->forced_variant
                 : (2 * p->uniform_tiles >= p->num_tiles ? kVariantPipelined : kVariantWarpStream);
  // the hub variant is offered to matrices that would take the warp-stream kernel, on
  // request (spblas_b200_plan_set_hub / SPBLAS_B200_HUB=1), for inspected plans only: the
  // no-info overloads re-derive the structure on every call and cannot pay for the analysis
  // ... with the table in shared memory while x fits in L2 (the bound is the L1 port), in
  // global memory when it does not (the bound is DRAM sectors)
  if (!forced && v == kVariantWarpStream && p->hub_enable && !p->light_inspect)
    v = double(p->csr_cols) * double(type_size_val(val_type)) > 0.75 * double(p->l2_bytes)
            ? kVariantHubGlobal
            : kVariantHubStream;
  if (!spmv_vec_ok(p, values))
    v = kVariantMergeTile;
  if (v != kVariantMergeTile && v != kVariantPipelined && v != kVariantWarpStream &&
      v != kVariantHubStream && v != kVariantHubGlobal)
    v = kVariantMergeTile;
  if ((v == kVariantHubStream || v == kVariantHubGlobal) &&
      (p->idx_type != SPBLAS_B200_I32 || p->host_exec_active))
    v = kVariantWarpStream; // a negative index marks a hub; chunked launches reload the table
  if (p->num_tiles > int64_t(0x7fffffff))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "too many tiles for one launch");
  if (v == kVariantHubStream || v == kVariantHubGlobal) {
    const bool global = v == kVariantHubGlobal;
    const int64_t cap = global ? hub_global_capacity(p, type_size_val(val_type))
                               : hub_capacity(p, type_size_val(val_type), kHubWarps);
    if (p->hub_state == 0 ||
        (p->hub_state == 1 && (p->hub_cap != cap || p->hub_by_popularity != global))) {
      if (int rc = build_hub_table(p, cap, global))
        return rc;
      // not worth a CTA-wide table unless a fair share of the gathers leaves the L2 port
      if (!forced && 20 * p->hub_refs < 3 * p->nnz) {
        release(p->hub_colind);
        release(p->hub_cols);
        p->hub_state = -1;
      }
    }
    if (p->hub_state != 1)
      v = kVariantWarpStream;
    else if (global)
      if (int rc = reserve(p, p->hub_x,
                           size_t(std::max<int64_t>(p->hub_count, 1)) * type_size_val(val_type)))
        return rc;
  }
  const bool ws = v == kVariantWarpStream || v == kVariantHubStream || v == kVariantHubGlobal;
  if (ws && p->ws_streams < 0) {
    const bool wide = type_size_val(val_type) == 8 || p->idx_type == SPBLAS_B200_I64;
    const int64_t resident =
        v == kVariantHubStream
            ? int64_t(p->num_sms) * kHubWarps
            : int64_t(p->num_sms) * (wide ? kWsCtasPerSmWide : kWsCtasPerSm) * kWsWarps;
    if (int rc = build_ws_partition(p, resident))
      return rc;
  }
  p->spmv_variant = v;
  *variant = v;
  if (starts)
    *starts = static_cast<const int64_t*>(ws ? p->ws_starts.p : p->tile_starts.p);
  if (units)
    *units = ws ? p->ws_streams : p->num_tiles;
  return SPBLAS_B200_SUCCESS;
}

int run_spmv_tiles(spblas_b200_plan* p, int val_type, const void* alpha,
                   const void* values, const void* x, void* y, int64_t T0,
                   int64_t T1) {
  int variant = 0;
  int64_t units = 0;
  if (int rc = prepare_spmv(p, val_type, values, &variant, nullptr, &units))
    return rc;
  if (T1 < 0) { // the whole product
    T0 = 0;
    T1 = units;
  }
  if (p->epi_d != nullptr)
    return spmv_dispatch_addend(p, val_type, variant, alpha, values, x, y, T0, T1);
  switch (val_type) {
  case SPBLAS_B200_F32:
    return dispatch_index<float, false>(p, variant, alpha, values, x, y, T0, T1);
  case SPBLAS_B200_F64:
    return dispatch_index<double, false>(p, variant, alpha, values, x, y, T0, T1);
  case SPBLAS_B200_S32:
    return dispatch_index<int32_t, false>(p, variant, alpha, values, x, y, T0, T1);
  default:
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "unknown value type");
  }
}

int run_spmv(spblas_b200_plan* p, int val_type, const void* alpha,
             const void* values, const void* x, void* y) {
  p->last_launches = 0;
  // a fused barrier of an earlier execute gave up on a peer: that step's x was incomplete.
  // (host-mapped flag: read without synchronising; cleared so that the caller can recover)
  if (p->barrier_gave_up_h && *static_cast<volatile unsigned int*>(p->barrier_gave_up_h)) {
    *static_cast<volatile unsigned int*>(p->barrier_gave_up_h) = 0u;
    p->barrier_gave_up_seen = true;
    return fail(p, SPBLAS_B200_CUDA_ERROR,
                "fused exchange: a peer did not reach the barrier of an earlier step within "
                "the timeout (SPBLAS_B200_BARRIER_TIMEOUT_MS); x of that step was incomplete");
  }
  return run_spmv_tiles(p, val_type, alpha, values, x, y, 0, -1);
}

} // namespace b200
