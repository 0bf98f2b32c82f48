// cabi.cu — the extern "C" surface declared in include/spblas_b200.h.
// Argument validation, plan lifetime, error strings and metadata queries.  No
// compute here; kernels live in inspect.cu / spmv.cu / spmm.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "plan.hpp"

namespace b200 {

int fail(spblas_b200_plan* p, int status, const std::string& msg) {
  if (p)
    p->err = msg;
  return status;
}

int cuda_fail(spblas_b200_plan* p, cudaError_t e, const char* what) {
  std::string msg = "CUDA encountered an error ";
  msg += cudaGetErrorName(e);
  msg += ": \"";
  msg += cudaGetErrorString(e);
  msg += "\" in ";
  msg += what;
  if (p)
    p->err = msg;
  return e == cudaErrorMemoryAllocation ? SPBLAS_B200_ALLOC_FAILED
                                        : SPBLAS_B200_CUDA_ERROR;
}

void release(DeviceBuffer& b) {
  if (b.p)
    cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

int reserve(spblas_b200_plan* p, DeviceBuffer& b, size_t bytes) {
  if (bytes == 0)
    bytes = 16;
  if (b.cap >= bytes)
    return SPBLAS_B200_SUCCESS;
  release(b);
  // grow geometrically so that the cached one-shot plan settles quickly
  size_t want = bytes + bytes / 4;
  want = (want + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    (void)cudaGetLastError();
    return fail(p, SPBLAS_B200_ALLOC_FAILED, "device allocation failed");
  }
  b.cap = want;
  return SPBLAS_B200_SUCCESS;
}

namespace {

void release_all(spblas_b200_plan* p) {
  DeviceBuffer* bufs[] = {&p->own_rowptr,  &p->own_colind, &p->own_perm,
                          &p->sort_tmp0,   &p->sort_tmp1,  &p->sort_tmp2,
                          &p->sort_ws,     &p->tile_starts, &p->tile_uniform, &p->carry_row,
                          &p->carry_val,   &p->segments,   &p->seg_partial,
                          &p->seg_counter, &p->stats,      &p->spmm_starts,
                          &p->spmm_carry_row, &p->spmm_carry_val, &p->barrier_state,
                          &p->ws_starts, &p->ws_carry_row, &p->ws_carry_val, &p->own_values,
                          &p->trsv_level, &p->trsv_order, &p->trsv_tmp0, &p->trsv_tmp1,
                          &p->trsv_level_ptr, &p->hub_cols, &p->hub_colind, &p->hub_x};
  for (DeviceBuffer* b : bufs)
    release(*b);
  release(p->hc_colmax);
  if (p->fp_status_h)
    cudaFreeHost(p->fp_status_h);
  p->fp_status_h = p->fp_status_d = nullptr;
  release(p->fp_state);
  if (p->barrier_gave_up_h)
    cudaFreeHost(p->barrier_gave_up_h);
  p->barrier_gave_up_h = p->barrier_gave_up_d = nullptr;
  release_host_exec(p);
  release_trsv_graphs(p);
  release(p->trsv_params);
  if (p->trsv_done_event)
    cudaEventDestroy(p->trsv_done_event);
  p->trsv_done_event = nullptr;
  if (p->trsv_capture_stream)
    cudaStreamDestroy(p->trsv_capture_stream);
  p->trsv_capture_stream = nullptr;
}

bool valid_index_type(int t) { return t == SPBLAS_B200_I32 || t == SPBLAS_B200_I64; }
bool valid_value_type(int t) {
  return t == SPBLAS_B200_F32 || t == SPBLAS_B200_F64 || t == SPBLAS_B200_S32;
}

// While alive, executes run on the plan's cached (image-order) values instead of gathering
// the caller's through the permutation — if the caller passed the pointer and type the
// cache was built from.
struct CachedValues {
  spblas_b200_plan* p;
  const void* saved_perm;
  bool active;
  CachedValues(spblas_b200_plan* plan, int val_type, const void*& values) : p(plan) {
    active = p->cached_values && p->csr_perm != nullptr && values == p->cached_src &&
             val_type == p->cached_val_type;
    saved_perm = p->csr_perm;
    if (active) {
      p->csr_perm = nullptr;
      values = p->own_values.p;
    }
  }
  ~CachedValues() { p->csr_perm = saved_perm; }
};

// NVTX range around every entry point that launches work (the reference has
// log_trace only).  Header-only NVTX3: a no-op costing nanoseconds unless a tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

thread_local std::string g_once_error;
// The one-shot plan of the overloads that take no operation_info_t, one per host thread.  It
// keeps the LAST structure's plan and its key; a call with the same key runs on the cached
// plan after a device-side check that the offsets array is unchanged (inspect.cu 1b).
struct OnceKey {
  int format = -1, off_type = 0, idx_type = 0;
  int64_t m = 0, n = 0, nnz = 0;
  const void* ptr = nullptr;
  const void* ind = nullptr;
  bool operator==(const OnceKey& o) const {
    return format == o.format && off_type == o.off_type && idx_type == o.idx_type && m == o.m &&
           n == o.n && nnz == o.nnz && ptr == o.ptr && ind == o.ind;
  }
};
constexpr int kOnceEntries = 8; // structures a thread's one-shot cache keeps (LRU)
struct OnceEntry {
  spblas_b200_plan* plan = nullptr;
  bool cached = false; // plan holds the light-inspected structure of `key`
  OnceKey key;
  unsigned long long last_use = 0;
};
struct OncePlanHolder {
  std::vector<OnceEntry> entries;
  unsigned long long clock = 0;
  void release_all_entries() {
    for (OnceEntry& e : entries)
      if (e.plan)
        spblas_b200_plan_destroy(e.plan);
    entries.clear();
  }
  ~OncePlanHolder() {
    // Thread exit: give the device buffers back while the runtime is still there (at process
    // exit it may already be unloading: then the context takes everything with it).
    if (!entries.empty() && cudaFree(nullptr) == cudaSuccess)
      release_all_entries();
    entries.clear();
  }
};
thread_local OncePlanHolder g_once;

} // namespace
} // namespace b200

using namespace b200;

extern "C" {

int spblas_b200_version(void) { return SPBLAS_B200_VERSION; }

const char* spblas_b200_status_string(int status) {
  switch (status) {
  case SPBLAS_B200_SUCCESS:
    return "success";
  case SPBLAS_B200_INVALID_ARGUMENT:
    return "invalid argument";
  case SPBLAS_B200_SHAPE_MISMATCH:
    return "shape mismatch";
  case SPBLAS_B200_NOT_SUPPORTED:
    return "not supported";
  case SPBLAS_B200_ALLOC_FAILED:
    return "allocation failed";
  case SPBLAS_B200_CUDA_ERROR:
    return "CUDA error";
  case SPBLAS_B200_NOT_INSPECTED:
    return "plan holds no inspected structure";
  case SPBLAS_B200_INVALID_STRUCTURE:
    return "invalid sparse structure";
  default:
    return "unknown status";
  }
}

int spblas_b200_plan_create(spblas_b200_plan** out, void* cuda_stream) {
  if (!out)
    return SPBLAS_B200_INVALID_ARGUMENT;
  *out = nullptr;
  spblas_b200_plan* p = new (std::nothrow) spblas_b200_plan();
  if (!p)
    return SPBLAS_B200_ALLOC_FAILED;
  p->stream = static_cast<cudaStream_t>(cuda_stream);
  cudaError_t e = cudaGetDevice(&p->device);
  if (e != cudaSuccess) {
    delete p;
    return SPBLAS_B200_CUDA_ERROR;
  }
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
  if (e != cudaSuccess) {
    delete p;
    return SPBLAS_B200_CUDA_ERROR;
  }
  p->num_sms = sms > 0 ? sms : 148;
  int l2 = 0;
  if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, p->device) == cudaSuccess && l2 > 0)
    p->l2_bytes = l2;
  int smem_sm = 0;
  if (cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, p->device) ==
          cudaSuccess && smem_sm > 0)
    p->smem_per_sm = size_t(smem_sm);
  if (const char* v = std::getenv("SPBLAS_B200_SPMV_VARIANT"))
    p->forced_variant = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_TILE_ITEMS")) {
    const int t = std::atoi(v);
    if (t >= 256 && t <= kSpmvMaxTileItems && t % 4 == 0)
      p->tile_items_override = t;
  }
  if (const char* v = std::getenv("SPBLAS_B200_STAGES"))
    p->stages = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_CTAS_PER_SM"))
    p->ctas_per_sm = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_WS_ITEMS")) {
    const int t = std::atoi(v);
    if (t >= 256 && t <= (1 << 20))
      p->ws_items_override = t;
  }
  if (const char* v = std::getenv("SPBLAS_B200_WS_CARVEOUT"))
    p->ws_carveout = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_HUB"))
    p->hub_enable = std::atoi(v) != 0;
  if (const char* v = std::getenv("SPBLAS_B200_HUB_COLS"))
    p->hub_cap_override = std::max<long long>(0, std::atoll(v));
  if (const char* v = std::getenv("SPBLAS_B200_HUB_MIN_COUNT"))
    p->hub_min_count = std::max<long long>(0, std::atoll(v));
  if (const char* v = std::getenv("SPBLAS_B200_TRSV_INSPECT"))
    p->trsv_relax_inspect = std::string(v) == "relax";
  if (const char* v = std::getenv("SPBLAS_B200_TRSV_GRAPH"))
    p->trsv_use_graph = std::atoi(v) != 0;
  if (const char* v = std::getenv("SPBLAS_B200_BARRIER_TIMEOUT_MS")) {
    const long long t = std::atoll(v);
    if (t > 0)
      p->barrier_timeout_ms = (unsigned long long)t;
  }
  if (const char* v = std::getenv("SPBLAS_B200_LATE_PUSH_ROWS"))
    p->late_push_max_rows = std::max<long long>(0, std::atoll(v));
  if (const char* v = std::getenv("SPBLAS_B200_HOST_CHUNKS"))
    p->host_chunks_override = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_SPMM_VARIANT"))
    p->spmm_forced = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_SPMM_CTAS_PER_SM"))
    p->spmm_ctas_per_sm = std::atoi(v);
  if (const char* v = std::getenv("SPBLAS_B200_SPMM_L2FRAC"))
    p->spmm_l2_fraction = float(std::atof(v));
  *out = p;
  return SPBLAS_B200_SUCCESS;
}

void spblas_b200_plan_destroy(spblas_b200_plan* p) {
  if (!p)
    return;
  release_all(p);
  delete p;
}

int spblas_b200_plan_set_stream(spblas_b200_plan* p, void* cuda_stream) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->stream = static_cast<cudaStream_t>(cuda_stream);
  return SPBLAS_B200_SUCCESS;
}

int spblas_b200_plan_set_scatter(spblas_b200_plan* p, int n_dst, void* const* d_dst,
                                 const int64_t* row_begin, const int64_t* row_end,
                                 int multicast) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->err.clear();
  if (n_dst < 0 || n_dst > kMaxPeers)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "at most SPBLAS_B200_MAX_PEERS destinations");
  if (n_dst > 0 && (!d_dst || !row_begin || !row_end))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null destination arrays");
  for (int d = 0; d < n_dst; ++d)
    if (!d_dst[d] || row_begin[d] < 0 || row_end[d] < row_begin[d])
      return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "bad scatter destination");
  p->scatter = ScatterSpec();
  p->scatter.n = n_dst;
  p->scatter.multicast = multicast ? 1 : 0;
  for (int d = 0; d < n_dst; ++d) {
    p->scatter.dst[d] = d_dst[d];
    p->scatter.lo[d] = row_begin[d];
    p->scatter.hi[d] = row_end[d];
  }
  return SPBLAS_B200_SUCCESS;
}

int spblas_b200_plan_set_barrier(spblas_b200_plan* p, int n_peers,
                                 void* const* d_remote_slots,
                                 const void* const* d_local_slots) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->err.clear();
  if (n_peers < 0 || n_peers > kMaxPeers)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "at most SPBLAS_B200_MAX_PEERS peers");
  if (n_peers > 0 && (!d_remote_slots || !d_local_slots))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null flag arrays");
  p->barrier = BarrierSpec();
  if (n_peers == 0)
    return SPBLAS_B200_SUCCESS;
  for (int d = 0; d < n_peers; ++d) {
    if (!d_remote_slots[d] || !d_local_slots[d])
      return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null flag slot");
    p->barrier.remote[d] = static_cast<unsigned long long*>(d_remote_slots[d]);
    p->barrier.local[d] = static_cast<const unsigned long long*>(d_local_slots[d]);
  }
  if (!p->barrier_state.p) {
    if (int rc = reserve(p, p->barrier_state, 2 * sizeof(unsigned int)))
      return rc;
    B200_CUDA_TRY(p, cudaMemsetAsync(p->barrier_state.p, 0, 2 * sizeof(unsigned int), p->stream));
  }
  if (!p->barrier_gave_up_h) {
    void* h = nullptr;
    B200_CUDA_TRY(p, cudaHostAlloc(&h, sizeof(unsigned int), cudaHostAllocMapped));
    p->barrier_gave_up_h = static_cast<unsigned int*>(h);
    *p->barrier_gave_up_h = 0u;
    void* d = nullptr;
    B200_CUDA_TRY(p, cudaHostGetDevicePointer(&d, h, 0));
    p->barrier_gave_up_d = static_cast<unsigned int*>(d);
  }
  p->barrier.n = n_peers;
  return SPBLAS_B200_SUCCESS;
}

int spblas_b200_plan_force_variant(spblas_b200_plan* p, int v) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->forced_variant = v;
  return SPBLAS_B200_SUCCESS;
}

int spblas_b200_plan_set_hub(spblas_b200_plan* p, int enable, int64_t max_cols,
                             int64_t min_count) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->err.clear();
  if (max_cols < -1 || min_count < -1)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "bad hub parameter");
  p->hub_enable = enable ? 1 : 0;
  if (max_cols < 0)
    max_cols = p->hub_cap_override; // -1: keep
  if (min_count < 0)
    min_count = p->hub_min_count;
  if (p->hub_cap_override != max_cols || p->hub_min_count != min_count)
    p->hub_state = 0; // the table (if any) was built under other limits
  p->hub_cap_override = max_cols;
  p->hub_min_count = min_count;
  return SPBLAS_B200_SUCCESS;
}

int spblas_b200_inspect(spblas_b200_plan* p, int format, int64_t m, int64_t n,
                        int64_t nnz, const void* d_ptr, const void* d_ind,
                        int off_type, int idx_type, int64_t k_hint, int flags) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_inspect");
  p->err.clear();
  p->inspected = false;
  // the plan is about to describe a product: a triangular structure it held shares the
  // index types and the owned buffers, and must not survive (trsv after this needs its own
  // trsv_inspect)
  if (p->trsv_ready) {
    p->trsv_ready = false;
    release_trsv_graphs(p);
  }
  p->host_chunks = 0; // the chunk table belongs to the previous structure
  p->cached_values = false;
  p->hub_state = 0; // so does the hub table
  p->light_inspect = (flags & SPBLAS_B200_INSPECT_LIGHT) != 0;
  if (format != SPBLAS_B200_CSR && format != SPBLAS_B200_CSC)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "format must be CSR or CSC");
  if (!valid_index_type(off_type) || !valid_index_type(idx_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "index/offset type must be int32 or int64");
  if (m < 0 || n < 0 || nnz < 0)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "negative dimension");
  const int64_t majors = format == SPBLAS_B200_CSR ? m : n;
  if (majors > 0 && d_ptr == nullptr)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "offsets pointer is null");
  if (nnz > 0 && d_ind == nullptr)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "indices pointer is null");
  if (majors == 0 && nnz != 0)
    return fail(p, SPBLAS_B200_INVALID_STRUCTURE, "nnz != 0 with no rows/columns");
  const int64_t imax = idx_type == SPBLAS_B200_I32 ? 0x7fffffff : INT64_MAX;
  if (m > imax || n > imax)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "dimension exceeds index type");
  if (off_type == SPBLAS_B200_I32 && nnz > 0x7fffffff)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "nnz exceeds offset type");

  p->format = format;
  p->m = m;
  p->n = n;
  p->nnz = nnz;
  p->user_ptr = d_ptr;
  p->user_ind = d_ind;
  p->off_type = off_type;
  p->idx_type = idx_type;
  p->k_hint = k_hint < 1 ? 1 : k_hint;

  int rc = inspect_structure(p, flags);
  if (rc == SPBLAS_B200_SUCCESS)
    p->inspected = true;
  return rc;
}

int spblas_b200_spmv(spblas_b200_plan* p, int val_type, const void* alpha,
                     const void* d_values, const void* d_x, void* d_y) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_spmv");
  p->err.clear();
  if (!p->inspected)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "spmv called before inspect");
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (!alpha)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "alpha is null");
  if ((p->nnz > 0 && !d_values) || (p->n > 0 && p->nnz > 0 && !d_x) ||
      (p->m > 0 && !d_y))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null device pointer");
  CachedValues scope(p, val_type, d_values);
  return run_spmv(p, val_type, alpha, d_values, d_x, d_y);
}

namespace {
// the addend of one execute: set for the duration of the call only
struct AddendScope {
  spblas_b200_plan* p;
  AddendScope(spblas_b200_plan* plan, int val_type, const void* beta, const void* d_d,
              int64_t ldd)
      : p(plan) {
    p->epi_d = d_d;
    p->epi_ldd = ldd;
    std::memcpy(p->epi_beta, beta, b200::type_size_val(val_type));
  }
  ~AddendScope() {
    p->epi_d = nullptr;
    p->epi_ldd = 0;
  }
};

// beta == 0 is "no addend": d is not read at all (a NaN in it must not reach y, exactly as
// stale contents of y never do)
bool beta_is_zero(int val_type, const void* beta) {
  switch (val_type) {
  case SPBLAS_B200_F32:
    return *static_cast<const float*>(beta) == 0.0f;
  case SPBLAS_B200_F64:
    return *static_cast<const double*>(beta) == 0.0;
  default:
    return *static_cast<const int32_t*>(beta) == 0;
  }
}
} // namespace

int spblas_b200_spmv_axpby(spblas_b200_plan* p, int val_type, const void* alpha,
                           const void* d_values, const void* d_x, const void* beta,
                           const void* d_d, void* d_y) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->err.clear();
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (!beta)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "beta is null");
  if (beta_is_zero(val_type, beta))
    return spblas_b200_spmv(p, val_type, alpha, d_values, d_x, d_y);
  if (p->inspected && p->m > 0 && !d_d)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null device pointer (d)");
  AddendScope addend(p, val_type, beta, d_d, 0);
  return spblas_b200_spmv(p, val_type, alpha, d_values, d_x, d_y);
}

int spblas_b200_spmm_axpby(spblas_b200_plan* p, int val_type, const void* alpha,
                           const void* d_values, const void* d_B, int64_t ldb,
                           const void* beta, const void* d_D, int64_t ldd, void* d_C,
                           int64_t ldc, int64_t k) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  p->err.clear();
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (!beta)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "beta is null");
  if (beta_is_zero(val_type, beta))
    return spblas_b200_spmm(p, val_type, alpha, d_values, d_B, ldb, d_C, ldc, k);
  if (ldd < k)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "leading dimension smaller than k");
  if (p->inspected && p->m > 0 && k > 0 && !d_D)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null device pointer (D)");
  AddendScope addend(p, val_type, beta, d_D, ldd);
  return spblas_b200_spmm(p, val_type, alpha, d_values, d_B, ldb, d_C, ldc, k);
}

int spblas_b200_plan_cache_values(spblas_b200_plan* p, int val_type, const void* d_values) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_plan_cache_values");
  p->err.clear();
  p->cached_values = false;
  if (!d_values)
    return SPBLAS_B200_SUCCESS; // cache dropped
  if (!p->inspected)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "cache_values called before inspect");
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (p->csr_perm == nullptr || p->nnz == 0)
    return SPBLAS_B200_SUCCESS; // CSR operands are streamed in place: nothing to cache
  if (int rc = reserve(p, p->own_values, size_t(p->nnz) * type_size_val(val_type)))
    return rc;
  if (int rc = gather_permuted_values(p, val_type, d_values, p->own_values.p))
    return rc;
  p->cached_values = true;
  p->cached_val_type = val_type;
  p->cached_src = d_values;
  return SPBLAS_B200_SUCCESS;
}

int spblas_b200_trsv_inspect(spblas_b200_plan* p, int64_t m, int64_t nnz,
                             const void* d_rowptr, const void* d_colind, int off_type,
                             int idx_type, int upper, int unit_diagonal) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_trsv_inspect");
  p->err.clear();
  p->trsv_ready = false;
  if (!valid_index_type(off_type) || !valid_index_type(idx_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "index/offset type must be int32 or int64");
  if (m < 0 || nnz < 0)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "negative dimension");
  if (m > 0x7fffffff)
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "triangular_solve supports at most 2^31-1 rows");
  if (m > 0 && !d_rowptr)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "offsets pointer is null");
  if (nnz > 0 && !d_colind)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "indices pointer is null");
  p->off_type = off_type;
  p->idx_type = idx_type;
  p->trsv_upper = upper ? 1 : 0;
  p->trsv_unit = unit_diagonal ? 1 : 0;
  p->trsv_m = m;
  p->trsv_nnz = nnz;
  p->trsv_rowptr = d_rowptr;
  p->trsv_colind = d_colind;
  // the plan now describes a triangular solve, not a product
  p->inspected = false;
  const int rc = trsv_inspect(p, m, d_rowptr, d_colind, p->trsv_upper, p->trsv_unit);
  if (rc == SPBLAS_B200_SUCCESS)
    p->trsv_ready = true;
  return rc;
}

int spblas_b200_trsv(spblas_b200_plan* p, int val_type, const void* alpha_a,
                     const void* alpha_b, const void* d_values, const void* d_b,
                     void* d_x) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_trsv");
  p->err.clear();
  if (!p->trsv_ready)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "trsv called before trsv_inspect");
  if (p->trsv_m > 0 && (!d_b || !d_x || !d_values))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null device pointer");
  return trsv_solve(p, val_type, alpha_a, alpha_b, d_values, d_b, d_x);
}

int spblas_b200_transpose_inspect(spblas_b200_plan* p, int64_t m, int64_t n, int64_t nnz,
                                  const void* d_rowptr, const void* d_colind,
                                  int off_type, int idx_type) {
  // A (m x n, CSR) read column-major IS A^T (n x m) stored as CSC over the same arrays:
  // the CSC inspect builds the row-major image of A^T, i.e. its CSR structure
  return spblas_b200_inspect(p, SPBLAS_B200_CSC, n, m, nnz, d_rowptr, d_colind, off_type,
                             idx_type, 1, SPBLAS_B200_INSPECT_LIGHT);
}

int spblas_b200_transpose(spblas_b200_plan* p, int val_type, const void* d_values,
                          void* d_t_rowptr, void* d_t_colind, void* d_t_values) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_transpose");
  p->err.clear();
  if (!p->inspected || p->format != SPBLAS_B200_CSC)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "transpose called before transpose_inspect");
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (!d_t_rowptr || (p->nnz > 0 && (!d_values || !d_t_colind || !d_t_values)))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null device pointer");
  return run_transpose(p, val_type, d_values, d_t_rowptr, d_t_colind, d_t_values);
}

int spblas_b200_spmv_host(spblas_b200_plan* p, int val_type, const void* alpha,
                          const void* d_values, const void* h_x, void* h_y,
                          void* d_x, void* d_y) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_spmv_host");
  p->err.clear();
  if (!p->inspected)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "spmv_host called before inspect");
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (!alpha)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "alpha is null");
  if ((p->nnz > 0 && !d_values) || (p->n > 0 && p->nnz > 0 && (!h_x || !d_x)) ||
      (p->m > 0 && (!h_y || !d_y)))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null pointer");
  CachedValues scope(p, val_type, d_values);
  return run_spmv_host(p, val_type, alpha, d_values, h_x, h_y, d_x, d_y);
}

int spblas_b200_spmm(spblas_b200_plan* p, int val_type, const void* alpha,
                     const void* d_values, const void* d_B, int64_t ldb,
                     void* d_C, int64_t ldc, int64_t k) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  NvtxRange nvtx_range("spblas_b200_spmm");
  p->err.clear();
  if (!p->inspected)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "spmm called before inspect");
  if (!valid_value_type(val_type))
    return fail(p, SPBLAS_B200_NOT_SUPPORTED, "value type must be f32, f64 or s32");
  if (!alpha)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "alpha is null");
  if (k < 0)
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "negative k");
  if (ldb < k || ldc < k)
    return fail(p, SPBLAS_B200_SHAPE_MISMATCH, "leading dimension smaller than k");
  if (k > 0 && ((p->nnz > 0 && (!d_values || !d_B)) || (p->m > 0 && !d_C)))
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "null device pointer");
  CachedValues scope(p, val_type, d_values);
  return run_spmm(p, val_type, alpha, d_values, d_B, ldb, d_C, ldc, k);
}

// The entry of the calling thread's cache for `key`: the one that holds it (a hit: *out_entry
// ->cached stays true), else a free or the least recently used one (its plan's buffers are
// reused for the new structure).
static int once_plan(void* stream, const OnceKey& key, OnceEntry** out_entry) {
  int dev = 0;
  cudaGetDevice(&dev);
  OnceEntry* pick = nullptr;
  for (OnceEntry& e : g_once.entries) {
    if (e.plan && e.plan->device != dev)
      continue; // buffers of another device
    if (e.cached && e.key == key) {
      pick = &e;
      break;
    }
  }
  if (!pick) {
    if (int(g_once.entries.size()) < kOnceEntries) {
      g_once.entries.emplace_back();
      pick = &g_once.entries.back();
    } else {
      for (OnceEntry& e : g_once.entries)
        if (!pick || e.last_use < pick->last_use)
          pick = &e;
      pick->cached = false;
      if (pick->plan && pick->plan->device != dev) {
        spblas_b200_plan_destroy(pick->plan);
        pick->plan = nullptr;
      }
    }
  }
  if (!pick->plan) {
    const int rc = spblas_b200_plan_create(&pick->plan, stream);
    if (rc) {
      g_once_error = "could not create the one-shot plan";
      return rc;
    }
  }
  pick->last_use = ++g_once.clock;
  pick->plan->stream = static_cast<cudaStream_t>(stream);
  *out_entry = pick;
  return SPBLAS_B200_SUCCESS;
}

// Waits (without a stream synchronisation) for the verify kernel of call `seq` to report:
// 1 unchanged, 0 changed, -1 the stream failed or never ran it.
static int await_verdict(spblas_b200_plan* p, unsigned int seq) {
  volatile unsigned long long* st = p->fp_status_h;
  for (unsigned long long spins = 0;; ++spins) {
    const unsigned long long v = *st;
    if ((v >> 1) == seq)
      return int(v & 1ull);
    if ((spins & 0xfffffull) == 0xfffffull) { // every ~1M polls: is the stream still alive?
      const cudaError_t e = cudaStreamQuery(p->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady)
        return -1;
      if (e == cudaSuccess && ((*st) >> 1) != seq)
        return -1; // everything ran, the verdict never came (e.g. a captured stream)
    }
  }
}

static int spmv_once_impl(void* stream, int format, int64_t m, int64_t n, int64_t nnz,
                          const void* d_ptr, const void* d_ind, int off_type, int idx_type,
                          int val_type, const void* alpha, const void* d_values,
                          const void* d_x, const void* beta, const void* d_d, void* d_y) {
  g_once_error.clear();
  OnceKey key;
  key.format = format, key.off_type = off_type, key.idx_type = idx_type;
  key.m = m, key.n = n, key.nnz = nnz, key.ptr = d_ptr, key.ind = d_ind;
  OnceEntry* entry = nullptr;
  int rc = once_plan(stream, key, &entry);
  if (rc)
    return rc;
  spblas_b200_plan* p = entry->plan;
  auto execute = [&]() {
    return beta ? spblas_b200_spmv_axpby(p, val_type, alpha, d_values, d_x, beta, d_d, d_y)
                : spblas_b200_spmv(p, val_type, alpha, d_values, d_x, d_y);
  };
  // ---- a structure seen before, if the offsets array still holds what it held ---------------
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(p->stream, &cap);
  if (entry->cached && format == SPBLAS_B200_CSR && m > 0 && p->fp_status_h &&
      cap == cudaStreamCaptureStatusNone) {
    const unsigned int seq = ++p->fp_seq == 0 ? ++p->fp_seq : p->fp_seq;
    p->inspected = true;
    rc = verify_structure(p, seq);
    if (rc == SPBLAS_B200_SUCCESS) {
      // the product is launched at once, behind the check: its kernels return without
      // touching y unless the check opens the gate
      p->gate = structure_gate(p);
      p->gate_value = seq;
      rc = execute();
      p->gate = nullptr;
      const int verdict = rc == SPBLAS_B200_SUCCESS ? await_verdict(p, seq) : -1;
      p->inspected = false;
      if (rc == SPBLAS_B200_SUCCESS && verdict == 1)
        return SPBLAS_B200_SUCCESS;
    }
    p->inspected = false;
    entry->cached = false; // changed in place (or the check could not run): start over
  }
  // ---- a structure not seen before: light inspect (validates, leaves the fingerprint) --------
  entry->cached = false;
  if (!p->fp_status_h) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess) {
      p->fp_status_h = static_cast<unsigned long long*>(h);
      *p->fp_status_h = 0;
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess)
        p->fp_status_d = static_cast<unsigned long long*>(d);
    }
  }
  rc = spblas_b200_inspect(p, format, m, n, nnz, d_ptr, d_ind, off_type, idx_type,
                           1, SPBLAS_B200_INSPECT_LIGHT);
  if (rc == SPBLAS_B200_SUCCESS)
    rc = execute();
  if (rc)
    g_once_error = p->err;
  else if (p->fp_status_d) {
    entry->cached = true; // a later call may reuse this plan — after the device-side check
    entry->key = key;
  }
  // the one-shot plan never trusts the caller's structure pointers between calls
  p->inspected = false;
  return rc;
}

int spblas_b200_spmv_once(void* stream, int format, int64_t m, int64_t n,
                          int64_t nnz, const void* d_ptr, const void* d_ind,
                          int off_type, int idx_type, int val_type,
                          const void* alpha, const void* d_values,
                          const void* d_x, void* d_y) {
  return spmv_once_impl(stream, format, m, n, nnz, d_ptr, d_ind, off_type, idx_type, val_type,
                        alpha, d_values, d_x, nullptr, nullptr, d_y);
}

int spblas_b200_spmv_axpby_once(void* stream, int format, int64_t m, int64_t n,
                                int64_t nnz, const void* d_ptr, const void* d_ind,
                                int off_type, int idx_type, int val_type,
                                const void* alpha, const void* d_values, const void* d_x,
                                const void* beta, const void* d_d, void* d_y) {
  if (!beta) {
    g_once_error = "beta is null";
    return SPBLAS_B200_INVALID_ARGUMENT;
  }
  return spmv_once_impl(stream, format, m, n, nnz, d_ptr, d_ind, off_type, idx_type, val_type,
                        alpha, d_values, d_x, beta, d_d, d_y);
}

void spblas_b200_once_release(void) { g_once.release_all_entries(); }

static int spmm_once_impl(void* stream, int format, int64_t m, int64_t n, int64_t nnz,
                          const void* d_ptr, const void* d_ind, int off_type, int idx_type,
                          int val_type, const void* alpha, const void* d_values,
                          const void* d_B, int64_t ldb, const void* beta, const void* d_D,
                          int64_t ldd, void* d_C, int64_t ldc, int64_t k) {
  g_once_error.clear();
  OnceKey key; // SpMM keeps no structure between calls: one entry of its own (format -2),
  key.format = -2; // whose buffers every SpMM one-shot call reuses
  OnceEntry* entry = nullptr;
  int rc = once_plan(stream, key, &entry);
  if (rc)
    return rc;
  spblas_b200_plan* p = entry->plan;
  entry->cached = true;
  entry->key = key;
  // SpMM wants to know about very long rows, so this is a full inspect.
  rc = spblas_b200_inspect(p, format, m, n, nnz, d_ptr, d_ind, off_type, idx_type,
                           k, SPBLAS_B200_INSPECT_DEFAULT);
  if (rc == SPBLAS_B200_SUCCESS)
    rc = beta ? spblas_b200_spmm_axpby(p, val_type, alpha, d_values, d_B, ldb, beta, d_D, ldd,
                                       d_C, ldc, k)
              : spblas_b200_spmm(p, val_type, alpha, d_values, d_B, ldb, d_C, ldc, k);
  if (rc)
    g_once_error = p->err;
  p->inspected = false;
  return rc;
}

int spblas_b200_spmm_once(void* stream, int format, int64_t m, int64_t n,
                          int64_t nnz, const void* d_ptr, const void* d_ind,
                          int off_type, int idx_type, int val_type,
                          const void* alpha, const void* d_values,
                          const void* d_B, int64_t ldb, void* d_C, int64_t ldc,
                          int64_t k) {
  return spmm_once_impl(stream, format, m, n, nnz, d_ptr, d_ind, off_type, idx_type, val_type,
                        alpha, d_values, d_B, ldb, nullptr, nullptr, 0, d_C, ldc, k);
}

int spblas_b200_spmm_axpby_once(void* stream, int format, int64_t m, int64_t n,
                                int64_t nnz, const void* d_ptr, const void* d_ind,
                                int off_type, int idx_type, int val_type,
                                const void* alpha, const void* d_values, const void* d_B,
                                int64_t ldb, const void* beta, const void* d_D, int64_t ldd,
                                void* d_C, int64_t ldc, int64_t k) {
  if (!beta) {
    g_once_error = "beta is null";
    return SPBLAS_B200_INVALID_ARGUMENT;
  }
  return spmm_once_impl(stream, format, m, n, nnz, d_ptr, d_ind, off_type, idx_type, val_type,
                        alpha, d_values, d_B, ldb, beta, d_D, ldd, d_C, ldc, k);
}

const char* spblas_b200_last_error(const spblas_b200_plan* p) {
  return p ? p->err.c_str() : "";
}

const char* spblas_b200_last_error_once(void) { return g_once_error.c_str(); }

int spblas_b200_plan_query(spblas_b200_plan* p, int what, void* out,
                           size_t bytes, size_t* needed) {
  if (!p)
    return SPBLAS_B200_INVALID_ARGUMENT;
  auto scalar = [&](int64_t v) -> int {
    if (needed)
      *needed = sizeof(int64_t);
    if (!out || bytes < sizeof(int64_t))
      return out ? fail(p, SPBLAS_B200_INVALID_ARGUMENT, "query buffer too small")
                 : SPBLAS_B200_SUCCESS;
    std::memcpy(out, &v, sizeof(v));
    return SPBLAS_B200_SUCCESS;
  };
  auto device_array = [&](const void* d, size_t n) -> int {
    if (needed)
      *needed = n;
    if (!out)
      return SPBLAS_B200_SUCCESS;
    if (bytes < n)
      return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "query buffer too small");
    if (n == 0)
      return SPBLAS_B200_SUCCESS;
    B200_CUDA_TRY(p, cudaMemcpyAsync(out, d, n, cudaMemcpyDeviceToHost, p->stream));
    B200_CUDA_TRY(p, cudaStreamSynchronize(p->stream));
    return SPBLAS_B200_SUCCESS;
  };
  switch (what) {
  case SPBLAS_B200_Q_BARRIER_EPOCH:
    return scalar(int64_t(p->barrier_epoch));
  case SPBLAS_B200_Q_BARRIER_TIMEOUT: {
    unsigned int st = 0;
    if (p->barrier_gave_up_h) {
      B200_CUDA_TRY(p, cudaStreamSynchronize(p->stream));
      st = *static_cast<volatile unsigned int*>(p->barrier_gave_up_h);
    }
    return scalar(int64_t(st != 0 || p->barrier_gave_up_seen));
  }
  case SPBLAS_B200_Q_TRSV_LEVELS:
    return scalar(p->trsv_ready ? p->trsv_levels : 0);
  case SPBLAS_B200_Q_TRSV_SWEEPS:
    return scalar(p->trsv_ready ? p->trsv_sweeps : 0);
  case SPBLAS_B200_Q_LAST_LAUNCHES:
    return scalar(p->last_launches);
  case SPBLAS_B200_Q_TOTAL_LAUNCHES:
    return scalar(p->total_launches);
  default:
    break;
  }
  if (!p->inspected)
    return fail(p, SPBLAS_B200_NOT_INSPECTED, "query on a plan with no structure");
  const size_t so = type_size_idx(p->off_type), si = type_size_idx(p->idx_type);
  switch (what) {
  case SPBLAS_B200_Q_NUM_TILES:
    return scalar(p->num_tiles);
  case SPBLAS_B200_Q_TILE_ITEMS:
    return scalar(p->tile_items);
  case SPBLAS_B200_Q_TILE_STARTS:
    return device_array(p->tile_starts.p,
                        size_t(p->num_tiles + 1) * 2 * sizeof(int64_t));
  case SPBLAS_B200_Q_TILE_UNIFORM:
    return device_array(p->tile_uniform.p, size_t(p->num_tiles) * sizeof(int));
  case SPBLAS_B200_Q_ROWLEN_HIST: {
    const size_t n = sizeof(int64_t) * SPBLAS_B200_HIST_BINS;
    if (needed)
      *needed = n;
    if (!out)
      return SPBLAS_B200_SUCCESS;
    if (bytes < n)
      return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "query buffer too small");
    if (!p->have_hist)
      return fail(p, SPBLAS_B200_NOT_INSPECTED, "histogram was skipped (light inspect)");
    std::memcpy(out, p->hist, n);
    return SPBLAS_B200_SUCCESS;
  }
  case SPBLAS_B200_Q_MAX_ROW_LEN:
    return scalar(p->max_row_len);
  case SPBLAS_B200_Q_EMPTY_ROWS:
    return scalar(p->empty_rows);
  case SPBLAS_B200_Q_SPMV_VARIANT:
    return scalar(p->spmv_variant);
  case SPBLAS_B200_Q_SPMM_VARIANT:
    return scalar(p->spmm_variant);
  case SPBLAS_B200_Q_CSR_ROWPTR:
    return device_array(p->csr_rowptr, size_t(p->csr_rows + 1) * so);
  case SPBLAS_B200_Q_CSR_COLIND:
    return device_array(static_cast<const char*>(p->csr_colind) +
                            (p->csr_perm ? 0 : size_t(p->base) * si),
                        size_t(p->nnz) * si);
  case SPBLAS_B200_Q_CSR_PERM:
    if (!p->csr_perm)
      return device_array(nullptr, 0);
    return device_array(p->csr_perm, size_t(p->nnz) * so);
  case SPBLAS_B200_Q_NUM_SEGMENTS:
    return scalar(p->num_segments);
  case SPBLAS_B200_Q_SEGMENTS:
    return device_array(p->segments.p, size_t(p->num_segments) * 3 * sizeof(int64_t));
  case SPBLAS_B200_Q_HUB_COUNT:
    return scalar(p->hub_state == 1 ? p->hub_count : 0);
  case SPBLAS_B200_Q_HUB_REFS:
    return scalar(p->hub_state == 1 ? p->hub_refs : 0);
  case SPBLAS_B200_Q_HUB_COLS:
    return device_array(p->hub_cols.p,
                        p->hub_state == 1 ? size_t(p->hub_count) * sizeof(int32_t) : 0);
  case SPBLAS_B200_Q_HUB_COLIND:
    if (p->hub_state != 1)
      return device_array(nullptr, 0);
    return device_array(static_cast<const char*>(p->hub_colind.p) +
                            size_t(p->base & 3) * sizeof(int32_t),
                        size_t(p->nnz) * sizeof(int32_t));
  default:
    return fail(p, SPBLAS_B200_INVALID_ARGUMENT, "unknown query selector");
  }
}

} // extern "C"
