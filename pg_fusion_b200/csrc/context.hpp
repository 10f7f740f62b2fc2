// Library context: device resources, scans (HBM-resident page sets), Bloom filter slots
// and join tables.  All CUDA calls go through CU() so errors become sticky status codes
// instead of aborting the host process.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

#include <cstdarg>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pgf_b200.h"
#include "device_types.cuh"

struct pgf_ctx;

namespace pgf {

// NVTX range over a scope (the reference traces planning / execution with `tracing`; here the phases of
// the hot path show up as named ranges in nsys / ncu timelines): ingest, fused pipelines, merges, exchanges.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Scan {
  uint64_t id = 0;
  std::vector<pgf_column_spec> schema;
  uint8_t* d_pages = nullptr;
  uint64_t cap_pages = 0;
  uint64_t npages = 0;
  uint64_t rows = 0;
  bool finished = false;
  bool descs_dirty = true;
  std::vector<PageDesc> h_descs;
  std::vector<LayoutClass> h_classes;
  PageDesc* d_descs = nullptr;
  uint64_t d_descs_cap = 0;
  LayoutClass* d_classes = nullptr;
  uint64_t d_classes_cap = 0;
  uint32_t max_page_rows = 0;   // largest row_count of any page
  uint64_t pending_async = 0;   // async copies in flight from caller-owned pinned pages
  // Recorded on the copy stream behind the last copy of THIS scan: finish and reset wait for it, not for the whole
  // copy stream -- the pages of several scans may be queued at once, and the first scan to finish must not wait
  // for the copies of the others.
  cudaEvent_t ev_pushed = nullptr;
  std::mutex mu;                // serialises producers of this scan
};

struct BloomSlot {
  pgf_bloom_params params{};
  uint64_t* d_words = nullptr;
  uint64_t lifecycle = 0;  // (generation << 2) | state, runtime_filter/src/shared.rs:7-9
  uint64_t set_bits = 0;   // popcount of the published bit array (saturation check of the fused probe)
  DevBloom dev{};
};

struct JoinTable {
  uint4* d_rows = nullptr;  // dense build rows (row sets: PGF_BUILD_ROWS_ONLY / PGF_XCHG_ROWS_ONLY); d_slots is null then
  size_t rows_alloc_bytes = 0;
  uint4* d_slots = nullptr;
  size_t alloc_bytes = 0;   // size of the allocation behind d_slots (recycled through pgf_ctx::join_cache)
  uint32_t capacity = 0;
  uint32_t slot_u4 = 1;
  uint64_t rows = 0;
  int32_t key_type = 0;
  uint32_t npayload = 0;
  int32_t payload_type[4] = {0, 0, 0, 0};
  uint8_t payload_nullable[4] = {1, 1, 1, 1};   // nullability of the build-side column behind each payload
  uint16_t payload_word[4] = {0, 0, 0, 0};
};

}  // namespace pgf

struct pgf_ctx {
  int device = 0;
  uint32_t page_size = 65536;
  uint32_t staging_pages = 512;
  uint32_t flags = 0;                   // PGF_CFG_*
  int sm_count = 148;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t compute_stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_copy = nullptr;
  // pinned staging (two chunks of staging_pages pages) for pageable caller memory
  uint8_t* staging[2] = {nullptr, nullptr};
  cudaEvent_t staging_ev[2] = {nullptr, nullptr};
  int staging_next = 0;
  pgf::Counters* d_counters = nullptr;
  pgf::Counters* h_counters = nullptr;  // pinned
  uint32_t* d_flags = nullptr;          // [0] group-table overflow, [1] table used, [2..] scratch
  uint32_t* h_flags = nullptr;          // pinned
  // Grow-only scratch of the aggregate pipelines: [header][group table][result entries], so a
  // run costs one memset, the fused kernel, the extract kernel and two small D2H copies
  // instead of a cudaMalloc/cudaFree cycle per table array.
  uint8_t* d_arena = nullptr;
  size_t d_arena_cap = 0;
  uint8_t* d_out = nullptr;             // result entries of tables too large for the arena prefix
  size_t d_out_cap = 0;
  uint8_t* d_topk = nullptr;            // scratch of the device top-k selection
  size_t d_topk_cap = 0;
  uint8_t* d_entries = nullptr;         // stage-C entries of a split join pipeline (grow-only)
  size_t d_entries_cap = 0;
  uint8_t* h_arena = nullptr;           // pinned mirror of the header and the first result entries
  pgf_runtime_filter_metrics rf_metrics{};   // RuntimeFilter* counters (runtime_metrics/src/lib.rs:125-131)
  bool partial_pending = false;         // an asynchronous partial run awaits its merge
  // Recycled join-table allocations: cudaMalloc / cudaFree of GB-sized tables cost 5-20 ms each
  // (page-table work, implicit synchronisation), more than the join kernels themselves.
  struct CachedBuf { void* p; size_t bytes; };
  std::vector<CachedBuf> join_cache;
  void* join_alloc(size_t bytes, size_t* got) {   // caller holds mu or is single threaded per the ABI contract
    size_t best = join_cache.size();
    for (size_t i = 0; i < join_cache.size(); ++i)
      if (join_cache[i].bytes >= bytes && join_cache[i].bytes <= 4 * bytes + (16u << 20) &&
          (best == join_cache.size() || join_cache[i].bytes < join_cache[best].bytes)) best = i;
    if (best != join_cache.size()) {
      void* p = join_cache[best].p;
      *got = join_cache[best].bytes;
      join_cache.erase(join_cache.begin() + long(best));
      return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      for (auto& c : join_cache) cudaFree(c.p);   // give the cache back and retry once
      join_cache.clear();
      if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    *got = bytes;
    return p;
  }
  void join_free(void* p, size_t bytes) {
    if (!p) return;
    if (join_cache.size() >= 16) {   // keep the sixteen most recent (a partitioned Q3 pass cycles through ~10 buffers)
      cudaFree(join_cache.front().p);
      join_cache.erase(join_cache.begin());
    }
    join_cache.push_back(CachedBuf{p, bytes});
  }
  // multi-GPU: NCCL communicator of this rank (comm.cpp) and grow-only exchange scratch
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  uint8_t* d_xchg = nullptr;
  size_t d_xchg_cap = 0;
  std::map<uint64_t, std::unique_ptr<pgf::Scan>> scans;
  std::map<uint64_t, pgf::BloomSlot> blooms;
  std::map<uint64_t, pgf::JoinTable> joins;
  uint64_t next_handle = 1;
  std::vector<void*> registered;
  std::mutex mu;
  mutable std::mutex err_mu;    // scan producers of different scans may fail concurrently
  std::string last_error;
  pgf_status sticky = PGF_OK;
  float last_kernel_ms = 0.f;

  pgf_status fail(pgf_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> g(err_mu);
    last_error = buf;
    return st;
  }
  pgf_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaErrorMemoryAllocation) {  // recoverable: the context stays usable
      cudaGetLastError();
      return fail(PGF_ERR_OUT_OF_MEMORY, "out of device memory in %s at %s:%d", what, file, line);
    }
    sticky = PGF_ERR_CUDA;
    return fail(PGF_ERR_CUDA, "CUDA error %d (%s) in %s at %s:%d", int(e), cudaGetErrorString(e), what, file, line);
  }
};

#define CU(ctx, call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define PGF_TRY(expr)                   \
  do {                                  \
    pgf_status st__ = (expr);           \
    if (st__ != PGF_OK) return st__;    \
  } while (0)

namespace pgf {
// internal entry points implemented in the .cu files
pgf_status scan_sync_descs(pgf_ctx* ctx, Scan& s);
pgf_status scan_device_validate(pgf_ctx* ctx, Scan& s);
pgf_status bloom_make_dev(const pgf_bloom_params& p, uint64_t* d_words, DevBloom* out);
pgf_status bloom_insert_host_keys(pgf_ctx* ctx, BloomSlot& b, const void* keys, int32_t key_width,
                                  const uint8_t* validity, uint64_t n, uint64_t* inserted);
pgf_status bloom_insert_scan(pgf_ctx* ctx, BloomSlot& b, Scan& s, uint32_t col, uint64_t* inserted);
pgf_status bloom_probe_host_keys(pgf_ctx* ctx, BloomSlot& b, bool ready, const void* keys,
                                 int32_t key_width, const uint8_t* validity, uint64_t n,
                                 uint8_t* decisions, pgf_probe_stats* stats);
pgf_status bloom_probe_scan(pgf_ctx* ctx, BloomSlot& b, bool ready, Scan& s, uint32_t col,
                            uint8_t* decisions, pgf_probe_stats* stats);
pgf_status bloom_or_strided(pgf_ctx* ctx, uint64_t* dst, const uint64_t* src, uint64_t nwords, uint32_t narrays, uint32_t grid);
pgf_status bloom_or_device(pgf_ctx* ctx, BloomSlot& b, const void* dev_words, uint64_t nwords,
                           uint32_t narrays);
pgf_status bloom_count_bits(pgf_ctx* ctx, BloomSlot& b);
pgf_status pipeline_run(pgf_ctx* ctx, const pgf_pipeline* plan, bool check_only, void* dev_state_out,
                        uint64_t state_cap, uint64_t* state_bytes, bool partial, pgf_result** out);
pgf_status pipeline_run_partial_async(pgf_ctx* ctx, const pgf_pipeline* plan, void* dev_state_out, uint64_t state_cap);
pgf_status pipeline_merge(pgf_ctx* ctx, const pgf_pipeline* plan, const void* dev_states,
                          uint64_t stride, uint32_t nstates, bool bounded, pgf_result** out);
pgf_status join_export(pgf_ctx* ctx, const JoinTable& jt, void* dev_rows_out, uint64_t capacity_rows, uint64_t* rows_out);
pgf_status join_from_fragments(pgf_ctx* ctx, const JoinTable& like, const void* dev_rows, uint64_t stride_bytes,
                               const uint64_t* counts, uint32_t nfragments, uint64_t* table_out);
pgf_status pipeline_run_sharded(pgf_ctx* ctx, const pgf_pipeline* plan, uint64_t max_groups, pgf_result** out);
pgf_status join_exchange(pgf_ctx* ctx, uint64_t handle, uint32_t mode, uint64_t* out_handle, uint64_t* nvlink_bytes);
pgf_status comm_all_gather(pgf_ctx* ctx, const void* send, void* recv, uint64_t bytes);   // compute stream, not synchronised
pgf_status comm_all_to_all_v(pgf_ctx* ctx, const void* send, const uint64_t* send_off, const uint64_t* send_bytes, void* recv,
                             const uint64_t* recv_off, const uint64_t* recv_bytes);
void comm_release(pgf_ctx* ctx);
pgf_status gen_scan(pgf_ctx* ctx, uint64_t scan_id, const pgf_gen_spec* spec);
pgf_status gen_schema(int32_t table, pgf_column_spec* schema, uint32_t* ncols);
}  // namespace pgf
