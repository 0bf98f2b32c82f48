// plan.hpp — the inspect-phase state behind `spblas_b200_plan` (opaque in the C ABI).
//
// It plays the role of __cusparse::spmv_state_t
// (reference: include/spblas/vendor/cusparse/detail/spmv_state_t.hpp:11-52) but
// owns real metadata: the merge-path partition table, the row-length histogram,
// the CSR image of a CSC matrix and the carry workspace, so that execute never
// allocates (the reference's cuSPARSE wrapper mallocs per call,
// vendor/cusparse/spmv_impl.hpp:74-89).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/spblas_b200.h"

namespace b200 {

// ---- tiling constants shared by inspect and execute -------------------------
constexpr int kSpmvThreads = 256;
// merge items (row ends + nonzeros) per CTA tile
constexpr int kSpmvTileItems = 2048;
// SpMM: rows longer than this are cut into segments of this many nonzeros
constexpr int64_t kSpmmSegment = 4096;

// largest tile the kernels' shared-memory carve-up is sized for
constexpr int kSpmvMaxTileItems = 4096;

// SpMV kernel variants (reported by SPBLAS_B200_Q_SPMV_VARIANT)
enum SpmvVariant : int {
  kVariantAuto = -1,
  // one tile per CTA, register-staged loads, products through shared memory.
  // Handles any pointer alignment; the fallback when colind/values are not
  // 16-byte aligned.
  kVariantMergeTile = 0,
  // persistent CTAs, TMA bulk copies of colind/values (cp.async.bulk + mbarrier)
  // and cp.async of the row ends into a multi-stage shared-memory ring, one
  // producer warp + 8 consumer warps that reduce rows straight out of the ring.
  kVariantPipelined = 1,
  // autonomous warps over warp-granular merge-path streams, 256 nonzeros per step,
  // rows reduced out of a per-warp slab with __syncwarp() only: the general path for
  // matrices bound by random gathers of x (needs 16-byte aligned colind/values).
  kVariantWarpStream = 2,
  // the warp-stream walk with x at the most referenced ("hub") columns held in shared
  // memory: one CTA of 32 warps per SM, a plan-owned re-encoded copy of colind (hub.cu).
  // For matrices with skewed column popularity (R-MAT); int32 indices only.
  kVariantHubStream = 3,
  // the warp-stream walk with x at the most referenced columns gathered, once per product,
  // into a compact table in GLOBAL memory (half of L2, hottest entries first) through the same
  // re-encoded colind: for matrices whose x is larger than L2, where a gather that misses L2
  // costs a 32-byte DRAM sector (C5: R-MAT scale 27, x = 1.07 GB).  int32 indices only.
  kVariantHubGlobal = 4,
};

struct DeviceBuffer {
  void* p = nullptr;
  size_t cap = 0;
};

// ---- fused exchange of a row-block sharded iteration (y -> x) ---------------------
// Rows [lo, hi) of the block this plan multiplies are ALSO stored, by the SpMV kernels
// themselves, to dst[row]: peer GPUs' replicas of the next x, mapped into this
// process (NVLink peer memory), or one multicast address (NVLS) that reaches them all.
constexpr int kMaxPeers = 8;
struct ScatterSpec {
  int n = 0;
  int multicast = 0;
  void* dst[kMaxPeers] = {};
  int64_t lo[kMaxPeers] = {}, hi[kMaxPeers] = {};
};
// One flag word per peer in every rank's flag array: after its stores a rank writes the
// step number into its slot on every peer and waits for every peer's number in its own.
struct BarrierSpec {
  int n = 0;
  unsigned long long* remote[kMaxPeers] = {};
  const unsigned long long* local[kMaxPeers] = {};
};

} // namespace b200

struct spblas_b200_plan {
  cudaStream_t stream = nullptr;
  int device = 0;
  int num_sms = 148;
  int64_t l2_bytes = 126ll << 20;
  size_t smem_per_sm = 228u << 10;

  // ---- structure as given by the caller (not owned) -------------------------
  bool inspected = false;
  int format = SPBLAS_B200_CSR;
  int off_type = SPBLAS_B200_I32;
  int idx_type = SPBLAS_B200_I32;
  int64_t m = 0, n = 0, nnz = 0; // logical shape of A
  const void* user_ptr = nullptr;
  const void* user_ind = nullptr;
  int64_t k_hint = 1;

  // ---- effective CSR structure the kernels run on ---------------------------
  // CSR input: aliases user_ptr/user_ind, perm == nullptr.
  // CSC input: owned row-major image (csr_rows = m, csr_cols = n) + value permutation.
  const void* csr_rowptr = nullptr;
  const void* csr_colind = nullptr;
  const void* csr_perm = nullptr;
  int64_t csr_rows = 0, csr_cols = 0;
  int64_t base = 0; // rowptr[0] of the effective structure
  b200::DeviceBuffer own_rowptr, own_colind, own_perm, sort_tmp0, sort_tmp1,
      sort_tmp2, sort_ws;

  // ---- cached values of a CSC / transposed operand (spblas_b200_plan_cache_values) ---
  // values[perm[q]] gathered once into image order, so that executes stream them like a
  // CSR matrix's instead of gathering through the permutation.  Valid only for calls
  // that pass the same values pointer and type; dropped by the next inspect.
  bool cached_values = false;
  int cached_val_type = 0;
  const void* cached_src = nullptr;
  b200::DeviceBuffer own_values;

  // ---- triangular solve (trsv.cu): level sets of the inspected triangle ---------------
  bool trsv_ready = false;
  int trsv_upper = 0, trsv_unit = 0;
  int64_t trsv_m = 0, trsv_nnz = 0, trsv_levels = 0, trsv_sweeps = 0;
  const void* trsv_rowptr = nullptr;
  const void* trsv_colind = nullptr;
  b200::DeviceBuffer trsv_level;     // int32 per row
  b200::DeviceBuffer trsv_order;     // int32 row ids, ascending level
  b200::DeviceBuffer trsv_tmp0, trsv_tmp1;
  b200::DeviceBuffer trsv_level_ptr; // int64 offsets of the levels in trsv_order
  std::vector<int64_t> trsv_level_ptr_h;
  // the level launches captured as a CUDA graph (one per value width), replayed per solve
  bool trsv_use_graph = true;            // env SPBLAS_B200_TRSV_GRAPH=0 launches level by level
  bool trsv_relax_inspect = false;       // env SPBLAS_B200_TRSV_INSPECT=relax: relaxation sweeps instead of frontiers
  cudaGraphExec_t trsv_graph[2] = {nullptr, nullptr};
  cudaStream_t trsv_capture_stream = nullptr;
  b200::DeviceBuffer trsv_params;        // operands of the current solve, read by the graph's kernels
  cudaEvent_t trsv_done_event = nullptr; // end of the last graph replay (orders the next rewrite of trsv_params)

  // ---- merge-path partition --------------------------------------------------
  int tile_items = b200::kSpmvTileItems;
  int tile_items_override = 0; // env SPBLAS_B200_TILE_ITEMS (tuning)
  int stages = 0;              // env SPBLAS_B200_STAGES (0 = default)
  int ctas_per_sm = 0;         // env SPBLAS_B200_CTAS_PER_SM (0 = default)
  int64_t num_tiles = 0;
  int64_t uniform_tiles = 0;      // tiles whose complete rows share one length <= 8
  b200::DeviceBuffer tile_starts; // int64 (row, nnz) pairs, num_tiles + 1 entries
  b200::DeviceBuffer tile_uniform; // int32 per tile: common length of the tile's complete rows (0: mixed)
  b200::DeviceBuffer carry_row;   // int64 per tile
  b200::DeviceBuffer carry_val;   // 8 bytes per tile

  // ---- SpMM segments ---------------------------------------------------------
  int64_t num_segments = 0;        // 0: no row is split
  int64_t num_split_rows = 0;
  b200::DeviceBuffer segments;     // int64 (row, begin, end) triples
  b200::DeviceBuffer seg_partial;  // partial C rows of split rows (num_segments x k)
  b200::DeviceBuffer seg_counter;  // int64 scratch

  // ---- SpMM streams (spmm_ring_kernel) ------------------------------------------
  // The merged sequence (row ends ++ nonzeros) cut into `spmm_streams` equal runs,
  // one per resident warp; stream w covers [spmm_starts[w], spmm_starts[w+1]).
  int64_t spmm_streams = 0;          // 0: table not built
  b200::DeviceBuffer spmm_starts;    // int64 (row, nnz) pairs, spmm_streams + 1 entries
  b200::DeviceBuffer spmm_carry_row; // int64 per stream: row the stream's tail belongs to, or -1
  b200::DeviceBuffer spmm_carry_val; // spmm_streams x k partial C rows
  int spmm_forced = -1;              // env SPBLAS_B200_SPMM_VARIANT: 0 group kernel, 1 stream kernel
  int spmm_ctas_per_sm = 0;          // env SPBLAS_B200_SPMM_CTAS_PER_SM (0 = what fits)
  float spmm_l2_fraction = -1.f;     // env SPBLAS_B200_SPMM_L2FRAC: share of B kept evict_last (<0: auto)

  // ---- SpMV warp streams (spmv_warp_stream_kernel) ----------------------------------
  int64_t ws_streams = -1;         // -1: table not built for the current structure
  int ws_items = 0;                // merge items per stream
  int ws_items_override = 0;       // env SPBLAS_B200_WS_ITEMS (tuning)
  int ws_carveout = -1;            // shared-memory carve-out in percent (env SPBLAS_B200_WS_CARVEOUT; -1 default)
  b200::DeviceBuffer ws_starts;    // int64 (row, nnz) pairs, ws_streams + 1 entries
  b200::DeviceBuffer ws_carry_row; // int64 per stream
  b200::DeviceBuffer ws_carry_val; // 8 bytes per stream

  // ---- hub columns (spmv_hub_stream_kernel, hub.cu) ----------------------------------
  // The `hub_count` most referenced columns, ascending, and a copy of the effective
  // colind in which a reference to hub number s reads ~s.  Built on the first execute
  // that wants the hub variant (the table's size depends on the value width).
  int hub_state = 0;           // 0: not analysed for the current structure, 1: table built, -1: analysed, no hubs
  int hub_enable = 0;          // env SPBLAS_B200_HUB / spblas_b200_plan_set_hub: the automatic choice may pick the hub variant
  int64_t hub_cap = 0;         // capacity (columns) the table was built for
  int64_t hub_cap_override = 0; // env SPBLAS_B200_HUB_COLS / set_hub (0: what shared memory holds)
  int64_t hub_min_count = 0;   // env SPBLAS_B200_HUB_MIN_COUNT / set_hub (0: 2 x SM count)
  bool hub_by_popularity = false; // the table's order: false ascending column (shared-memory kernel), true descending popularity (global-memory kernel)
  b200::DeviceBuffer hub_x;    // global-memory kernel: x at the hub columns, refilled by every product
  int64_t hub_count = 0;       // H
  int64_t hub_refs = 0;        // nonzeros that reference a hub column
  bool light_inspect = false;  // the current structure came from a LIGHT inspect (no-info overloads)
  bool host_exec_active = false; // inside spblas_b200_spmv_host (chunked launches: no hub variant)
  b200::DeviceBuffer hub_cols;   // int32[H]
  b200::DeviceBuffer hub_colind; // int32[(base & 3) + nnz]

  // ---- statistics -------------------------------------------------------------
  bool have_hist = false;
  int64_t hist[SPBLAS_B200_HIST_BINS] = {0};
  int64_t max_row_len = 0;
  int64_t empty_rows = 0;
  b200::DeviceBuffer stats; // device scratch for hist/max/flags

  // ---- fused exchange (set by spblas_b200_plan_set_scatter / _set_barrier) ---------
  b200::ScatterSpec scatter;
  b200::BarrierSpec barrier;
  unsigned long long barrier_epoch = 0; // steps signalled so far
  b200::DeviceBuffer barrier_state;     // uint32 block counter
  // "a wait gave up" flag in host-mapped memory: the kernel sets it, the next execute reads it
  // without a synchronisation and fails (a timeout must never be silent)
  unsigned int* barrier_gave_up_h = nullptr;
  unsigned int* barrier_gave_up_d = nullptr;
  bool barrier_gave_up_seen = false;           // sticky copy for SPBLAS_B200_Q_BARRIER_TIMEOUT
  unsigned long long barrier_timeout_ms = 30000; // env SPBLAS_B200_BARRIER_TIMEOUT_MS
  // an exchange of at most this many rows is pushed by the fix-up kernel's last block instead
  // of being stored by the product kernel (env SPBLAS_B200_LATE_PUSH_ROWS; 0: never)
  int64_t late_push_max_rows = 32768;

  // ---- host-buffer execute (spblas_b200_spmv_host, host_exec.cu) -----------------
  // The tiles cut into `host_chunks` consecutive chunks; chunk c covers tiles
  // [hc_tile[c], hc_tile[c+1]), completes rows [hc_row[c], hc_row[c+1]) and reads
  // x[0 .. hc_xneed[c]) at most (running maximum of the columns referenced so far), so
  // the upload of x, the products and the download of y overlap chunk by chunk.
  int host_chunks = 0; // 0: table not built for the current structure
  int hc_variant = -1; // the kernel variant (hence partition) the table was cut for
  std::vector<int64_t> hc_tile, hc_row, hc_xneed;
  int64_t hc_xlo = 0; // smallest column referenced: x below it is never uploaded
  b200::DeviceBuffer hc_colmax; // int64 per chunk (device scratch of the table build)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> hc_events; // 2 per chunk + 2
  int host_chunks_override = 0;       // env SPBLAS_B200_HOST_CHUNKS

  // ---- structure cache of the one-shot plan (spblas_b200_spmv_once; inspect.cu 1b) -----
  b200::DeviceBuffer fp_state;                 // HashState: fingerprint of the offsets, gate
  unsigned long long* fp_status_h = nullptr;   // host-mapped: (seq << 1) | same, written by the verify kernel
  unsigned long long* fp_status_d = nullptr;
  unsigned int fp_seq = 0;                     // verified calls so far
  // the kernels of a call that runs on a cached structure check *gate == gate_value first
  // and return without touching y otherwise (nullptr: no gate)
  const unsigned int* gate = nullptr;
  unsigned int gate_value = 0;

  // ---- addend of the running execute (4-argument multiply: y = alpha A x + beta d) -------
  // set by the *_axpby entry points for the duration of one call; nullptr: beta = 0
  const void* epi_d = nullptr;
  int64_t epi_ldd = 0;                  // SpMM: leading dimension of D
  alignas(8) unsigned char epi_beta[8] = {0};

  int forced_variant = -1;
  int spmv_variant = b200::kVariantMergeTile;
  int spmm_variant = 0;
  int64_t last_launches = 0;
  int64_t total_launches = 0;

  std::string err;
};

namespace b200 {

// status helpers (cabi.cu)
int fail(spblas_b200_plan* p, int status, const std::string& msg);
int cuda_fail(spblas_b200_plan* p, cudaError_t e, const char* what);
int reserve(spblas_b200_plan* p, DeviceBuffer& b, size_t bytes);
void release(DeviceBuffer& b);

#define B200_CUDA_TRY(plan, expr)                                              \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess)                                                    \
      return ::b200::cuda_fail((plan), e__, #expr);                            \
  } while (0)

// inspect.cu
int inspect_structure(spblas_b200_plan* p, int flags);
int verify_structure(spblas_b200_plan* p, unsigned int seq);
const unsigned int* structure_gate(const spblas_b200_plan* p);
int build_stream_partition(spblas_b200_plan* p, int64_t streams);
int build_ws_partition(spblas_b200_plan* p, int64_t resident_warps);
int run_transpose(spblas_b200_plan* p, int val_type, const void* values, void* t_rowptr,
                  void* t_colind, void* t_values);
int build_column_structure(spblas_b200_plan* p, int64_t m, int64_t nnz, const void* d_rowptr,
                           const void* d_colind);
int gather_permuted_values(spblas_b200_plan* p, int val_type, const void* values, void* out);
// hub.cu
int64_t hub_capacity(const spblas_b200_plan* p, size_t val_bytes, int walk_warps);
int64_t hub_global_capacity(const spblas_b200_plan* p, size_t val_bytes);
int build_hub_table(spblas_b200_plan* p, int64_t cap, bool by_popularity);
// spmv.cu
int run_spmv(spblas_b200_plan* p, int val_type, const void* alpha,
             const void* values, const void* x, void* y);
int run_spmv_tiles(spblas_b200_plan* p, int val_type, const void* alpha,
                   const void* values, const void* x, void* y, int64_t tile_begin,
                   int64_t tile_end);
// the kernel variant this product will use and the partition it runs on (tiles, or warp
// streams); builds the warp-stream table on first use
int prepare_spmv(spblas_b200_plan* p, int val_type, const void* values, int* variant,
                 const int64_t** starts, int64_t* units);
// host_exec.cu
int run_spmv_host(spblas_b200_plan* p, int val_type, const void* alpha,
                  const void* values, const void* h_x, void* h_y, void* d_x,
                  void* d_y);
void release_host_exec(spblas_b200_plan* p);
// trsv.cu
int trsv_inspect(spblas_b200_plan* p, int64_t m, const void* d_rowptr, const void* d_colind,
                 int upper, int unit);
void release_trsv_graphs(spblas_b200_plan* p);
int trsv_solve(spblas_b200_plan* p, int val_type, const void* alpha_a, const void* alpha_b,
               const void* values, const void* b, void* x);
// spmm.cu
int run_spmm(spblas_b200_plan* p, int val_type, const void* alpha,
             const void* values, const void* B, int64_t ldb, void* C,
             int64_t ldc, int64_t k);

inline size_t type_size_idx(int t) { return t == SPBLAS_B200_I64 ? 8 : 4; }
inline size_t type_size_val(int t) { return t == SPBLAS_B200_F64 ? 8 : 4; }

} // namespace b200
