// extern "C" surface of libpgf_b200.so: context, host-side layout, scan ingest and the
// Bloom lifecycle.  Kernels live in the .cu files.
#include <algorithm>
#include <cstring>

#include "context.hpp"
#include "layout.hpp"

using namespace pgf;

namespace {

Scan* find_scan(pgf_ctx* ctx, uint64_t id) {
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->scans.find(id);
  return it == ctx->scans.end() ? nullptr : it->second.get();
}

bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

pgf_status scan_reserve(pgf_ctx* ctx, Scan& s, uint64_t want_pages) {
  if (want_pages <= s.cap_pages) return PGF_OK;
  uint64_t cap = s.cap_pages ? s.cap_pages : 64;
  while (cap < want_pages) cap *= 2;
  uint8_t* fresh = nullptr;
  cudaError_t e = cudaMalloc(&fresh, cap * uint64_t(ctx->page_size));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate %llu pages of HBM for scan %llu",
                     (unsigned long long)cap, (unsigned long long)s.id);
  }
  if (s.npages) {
    // order the move after every copy already queued for this scan
    CU(ctx, cudaMemcpyAsync(fresh, s.d_pages, s.npages * uint64_t(ctx->page_size), cudaMemcpyDeviceToDevice,
                            ctx->copy_stream));
    CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  }
  if (s.d_pages) CU(ctx, cudaFree(s.d_pages));
  s.d_pages = fresh;
  s.cap_pages = cap;
  return PGF_OK;
}

// Host part of the import checks + descriptor-table entry for one page.
pgf_status scan_admit_page(pgf_ctx* ctx, Scan& s, const uint8_t* page, uint32_t len) {
  if (len < kPageHeaderLen || len > ctx->page_size)
    return ctx->fail(PGF_ERR_IMPORT_PAGE_HEADER_INVALID, "page length %u outside [20, %u]", len, ctx->page_size);
  uint16_t kind, flags;
  uint32_t payload_len;
  if (pgf_status st = decode_page_header(page, &kind, &flags, &payload_len))
    return ctx->fail(st, "scan %llu page %llu: bad transfer page header", (unsigned long long)s.id,
                     (unsigned long long)s.npages);
  if (uint64_t(payload_len) + kPageHeaderLen > len)
    return ctx->fail(PGF_ERR_LAYOUT_BLOCK_SLICE_TOO_SMALL, "payload_len %u exceeds page length %u", payload_len, len);
  const uint8_t* block = page + kPageHeaderLen;
  if (pgf_status st = check_block_structure(kind, flags, block, payload_len, s.schema.data(),
                                            uint32_t(s.schema.size())))
    return ctx->fail(st, "scan %llu page %llu rejected by import checks (status %d)", (unsigned long long)s.id,
                     (unsigned long long)s.npages, int(st));
  BlockHeader h;
  std::memcpy(&h, block, sizeof h);
  // layout class = distinct max_rows (offsets are a pure function of schema and max_rows)
  uint32_t cls = 0;
  for (; cls < s.h_classes.size(); ++cls)
    if (s.h_classes[cls].max_rows == h.max_rows) break;
  if (cls == s.h_classes.size()) {
    if (cls >= 65535) return ctx->fail(PGF_ERR_UNSUPPORTED_DATA, "too many distinct page shapes in one scan");
    LayoutClass lc{};
    lc.max_rows = h.max_rows;
    lc.pool_base = h.pool_base + kPageHeaderLen;
    for (uint32_t c = 0; c < h.col_count; ++c) {
      ColumnDesc d;
      std::memcpy(&d, block + sizeof(BlockHeader) + size_t(c) * sizeof(ColumnDesc), sizeof d);
      lc.values_off[c] = d.values_off + kPageHeaderLen;
      lc.validity_off[c] = d.validity_off + kPageHeaderLen;
    }
    s.h_classes.push_back(lc);
  }
  PageDesc pd{};
  pd.row_count = h.row_count;
  pd.layout_class = uint16_t(cls);
  pd.row_base = s.rows;
  for (uint32_t c = 0; c < h.col_count; ++c) {
    ColumnDesc d;
    std::memcpy(&d, block + sizeof(BlockHeader) + size_t(c) * sizeof(ColumnDesc), sizeof d);
    if (d.null_count) pd.null_mask |= uint16_t(1u << c);
  }
  s.h_descs.push_back(pd);
  s.rows += h.row_count;
  if (h.row_count > s.max_page_rows) s.max_page_rows = h.row_count;
  s.descs_dirty = true;
  return PGF_OK;
}

// behind the last copy this call queued for the scan
pgf_status scan_mark_pushed(pgf_ctx* ctx, Scan& s) {
  if (!s.ev_pushed) CU(ctx, cudaEventCreateWithFlags(&s.ev_pushed, cudaEventDisableTiming));
  CU(ctx, cudaEventRecord(s.ev_pushed, ctx->copy_stream));
  return PGF_OK;
}

}  // namespace

extern "C" {

pgf_status pgf_device_count(int32_t* count_out) {
  if (!count_out) return PGF_ERR_INVALID_ARGUMENT;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count_out = n;
  return PGF_OK;
}

pgf_status pgf_ctx_create(const pgf_config* config, pgf_ctx** ctx_out) {
  if (!ctx_out) return PGF_ERR_INVALID_ARGUMENT;
  *ctx_out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return PGF_ERR_NO_DEVICE;  // there is deliberately no CPU fallback
  }
  pgf_config cfg{};
  if (config) cfg = *config;
  if (cfg.device < 0 || cfg.device >= n) return PGF_ERR_INVALID_ARGUMENT;
  auto* ctx = new (std::nothrow) pgf_ctx();
  if (!ctx) return PGF_ERR_OUT_OF_MEMORY;
  ctx->device = cfg.device;
  ctx->page_size = cfg.page_size ? cfg.page_size : 65536u;
  ctx->staging_pages = cfg.staging_pages ? cfg.staging_pages : 512u;
  ctx->flags = cfg.flags;
  auto bail = [&](pgf_status st) {
    pgf_ctx_destroy(ctx);
    return st;
  };
  if (ctx->page_size % 16 != 0 || ctx->page_size < 1024 || ctx->page_size > (1u << 20)) return bail(PGF_ERR_INVALID_ARGUMENT);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return bail(PGF_ERR_CUDA);
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return bail(PGF_ERR_CUDA);
  if (prop.major != 10 || prop.minor != 0) return bail(PGF_ERR_NO_DEVICE);  // the kernels are sm_100a code: only compute capability 10.0 runs them
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(PGF_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(PGF_ERR_CUDA);
  if (cudaEventCreate(&ctx->ev_a) != cudaSuccess || cudaEventCreate(&ctx->ev_b) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) != cudaSuccess)
    return bail(PGF_ERR_CUDA);
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&ctx->staging_ev[i], cudaEventDisableTiming) != cudaSuccess) return bail(PGF_ERR_CUDA);
  if (cudaMalloc(&ctx->d_counters, sizeof(Counters)) != cudaSuccess) return bail(PGF_ERR_OUT_OF_MEMORY);
  if (cudaMallocHost(&ctx->h_counters, sizeof(Counters)) != cudaSuccess) return bail(PGF_ERR_OUT_OF_MEMORY);
  if (cudaMalloc(&ctx->d_flags, 64 * sizeof(uint32_t)) != cudaSuccess) return bail(PGF_ERR_OUT_OF_MEMORY);
  if (cudaMallocHost(&ctx->h_flags, 64 * sizeof(uint32_t)) != cudaSuccess) return bail(PGF_ERR_OUT_OF_MEMORY);
  if (cudaMallocHost(&ctx->h_arena, 64 * 1024) != cudaSuccess) return bail(PGF_ERR_OUT_OF_MEMORY);
  *ctx_out = ctx;
  return PGF_OK;
}

void pgf_ctx_destroy(pgf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& kv : ctx->scans) {
    if (kv.second->ev_pushed) cudaEventDestroy(kv.second->ev_pushed);
    if (kv.second->d_pages) cudaFree(kv.second->d_pages);
    if (kv.second->d_descs) cudaFree(kv.second->d_descs);
    if (kv.second->d_classes) cudaFree(kv.second->d_classes);
  }
  for (auto& kv : ctx->blooms)
    if (kv.second.d_words) cudaFree(kv.second.d_words);
  comm_release(ctx);
  if (ctx->d_xchg) cudaFree(ctx->d_xchg);
  for (auto& kv : ctx->joins) {
    if (kv.second.d_slots) cudaFree(kv.second.d_slots);
    if (kv.second.d_rows) cudaFree(kv.second.d_rows);
  }
  for (auto& c : ctx->join_cache) cudaFree(c.p);
  for (void* p : ctx->registered) cudaHostUnregister(p);
  for (int i = 0; i < 2; ++i) {
    if (ctx->staging[i]) cudaFreeHost(ctx->staging[i]);
    if (ctx->staging_ev[i]) cudaEventDestroy(ctx->staging_ev[i]);
  }
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->d_flags) cudaFree(ctx->d_flags);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  if (ctx->d_arena) cudaFree(ctx->d_arena);
  if (ctx->d_out) cudaFree(ctx->d_out);
  if (ctx->d_topk) cudaFree(ctx->d_topk);
  if (ctx->d_entries) cudaFree(ctx->d_entries);
  if (ctx->h_arena) cudaFreeHost(ctx->h_arena);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
  cudaGetLastError();
  delete ctx;
}

const char* pgf_last_error(const pgf_ctx* ctx) {
  if (!ctx) return "null context";
  // a per-thread copy: the message stays valid for the caller while other threads keep failing
  thread_local std::string copy;
  std::lock_guard<std::mutex> g(ctx->err_mu);
  copy = ctx->last_error;
  return copy.c_str();
}

pgf_status pgf_ctx_register_host_region(pgf_ctx* ctx, void* base, size_t len) {
  if (!ctx || !base || !len) return PGF_ERR_INVALID_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaHostRegister(base, len, cudaHostRegisterPortable));
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->registered.push_back(base);
  return PGF_OK;
}

pgf_status pgf_ctx_unregister_host_region(pgf_ctx* ctx, void* base) {
  if (!ctx || !base) return PGF_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  for (size_t i = 0; i < ctx->registered.size(); ++i)
    if (ctx->registered[i] == base) {
      CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
      CU(ctx, cudaHostUnregister(base));
      ctx->registered.erase(ctx->registered.begin() + long(i));
      return PGF_OK;
    }
  return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "host region %p was not registered", base);
}

pgf_status pgf_ctx_synchronize(pgf_ctx* ctx) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  return PGF_OK;
}

float pgf_ctx_last_kernel_ms(const pgf_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }

pgf_status pgf_ctx_runtime_filter_metrics(pgf_ctx* ctx, pgf_runtime_filter_metrics* out) {
  if (!ctx || !out) return PGF_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  *out = ctx->rf_metrics;
  return PGF_OK;
}
pgf_status pgf_ctx_note_pool_exhausted(pgf_ctx* ctx) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  ctx->rf_metrics.pool_exhausted_total++;
  return PGF_OK;
}

void* pgf_ctx_compute_stream(pgf_ctx* ctx) { return ctx ? (void*)ctx->compute_stream : nullptr; }

/* ---- host-side layout ---- */
pgf_status pgf_layout_plan_new(const pgf_column_spec* specs, uint32_t ncols, uint32_t max_rows,
                               uint32_t block_size, pgf_layout_plan* plan_out) {
  if ((!specs && ncols) || !plan_out) return PGF_ERR_INVALID_ARGUMENT;
  return plan_layout(specs, ncols, max_rows, block_size, plan_out);
}
pgf_status pgf_layout_fixed_row_cap(const pgf_column_spec* specs, uint32_t ncols, uint32_t block_size,
                                    uint32_t* cap_out) {
  if ((!specs && ncols) || !cap_out) return PGF_ERR_INVALID_ARGUMENT;
  return fixed_row_cap(specs, ncols, block_size, cap_out);
}
pgf_status pgf_block_validate(const uint8_t* block, size_t len) {
  if (!block) return PGF_ERR_INVALID_ARGUMENT;
  return validate_block(block, len, /*allow_ext=*/false);
}
pgf_status pgf_block_validate_ext(const uint8_t* block, size_t len, uint32_t extensions) {
  if (!block || (extensions & ~uint32_t(PGF_LAYOUT_EXT_DECIMAL128))) return PGF_ERR_INVALID_ARGUMENT;
  return validate_block(block, len, (extensions & PGF_LAYOUT_EXT_DECIMAL128) != 0);
}
pgf_status pgf_block_import_check(uint16_t kind, uint16_t flags, const uint8_t* block, size_t len,
                                  const pgf_column_spec* schema, uint32_t ncols) {
  if (!block || (!schema && ncols)) return PGF_ERR_INVALID_ARGUMENT;
  return check_block_full(kind, flags, block, len, schema, ncols);
}
pgf_status pgf_block_init(uint8_t* block, size_t len, const pgf_layout_plan* plan) {
  if (!block || !plan) return PGF_ERR_INVALID_ARGUMENT;
  return init_block(block, len, *plan);
}
pgf_status pgf_block_write_column(uint8_t* block, size_t len, uint32_t col, uint32_t nrows, const void* values,
                                  const uint8_t* validity) {
  if (!block || (!values && nrows)) return PGF_ERR_INVALID_ARGUMENT;
  return write_column(block, len, col, nrows, values, validity);
}
pgf_status pgf_block_set_row_count(uint8_t* block, size_t len, uint32_t nrows) {
  if (!block) return PGF_ERR_INVALID_ARGUMENT;
  return set_row_count(block, len, nrows);
}
pgf_status pgf_page_header_encode(uint16_t kind, uint16_t flags, uint32_t payload_len, uint8_t out[20]) {
  if (!out) return PGF_ERR_INVALID_ARGUMENT;
  encode_page_header(kind, flags, payload_len, out);
  return PGF_OK;
}
pgf_status pgf_page_header_decode(const uint8_t in[20], uint16_t* kind, uint16_t* flags, uint32_t* payload_len) {
  if (!in || !kind || !flags || !payload_len) return PGF_ERR_INVALID_ARGUMENT;
  return decode_page_header(in, kind, flags, payload_len);
}

/* ---- scans ---- */
pgf_status pgf_scan_declare(pgf_ctx* ctx, uint64_t scan_id, const pgf_column_spec* schema, uint32_t ncols,
                            uint64_t expected_pages) {
  if (!ctx || (!schema && ncols)) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  if (ncols > PGF_MAX_COLS) return ctx->fail(PGF_ERR_NOT_ELIGIBLE, "scan has %u columns; at most %u are staged", ncols, PGF_MAX_COLS);
  for (uint32_t c = 0; c < ncols; ++c)
    if (!known_type(schema[c].type_tag)) return ctx->fail(PGF_ERR_LAYOUT_INVALID_TYPE_TAG, "column %u has unknown type tag %u", c, schema[c].type_tag);
  CU(ctx, cudaSetDevice(ctx->device));
  auto s = std::make_unique<Scan>();
  s->id = scan_id;
  s->schema.assign(schema, schema + ncols);
  if (expected_pages) PGF_TRY(scan_reserve(ctx, *s, expected_pages));
  std::lock_guard<std::mutex> g(ctx->mu);
  if (ctx->scans.count(scan_id)) {
    if (s->d_pages) cudaFree(s->d_pages);
    return ctx->fail(PGF_ERR_STATE, "scan %llu already declared", (unsigned long long)scan_id);
  }
  ctx->scans[scan_id] = std::move(s);
  return PGF_OK;
}

pgf_status pgf_scan_push_pages(pgf_ctx* ctx, uint64_t scan_id, const uint8_t* pages, uint64_t npages,
                               uint64_t stride) {
  NvtxRange nvtx_("pgf:scan_push_pages");
  if (!ctx || (!pages && npages)) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  Scan* sp = find_scan(ctx, scan_id);
  if (!sp) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  Scan& s = *sp;
  std::lock_guard<std::mutex> g(s.mu);
  if (s.finished) return ctx->fail(PGF_ERR_STATE, "scan %llu already finished", (unsigned long long)scan_id);
  if (npages == 0) return PGF_OK;
  if (stride < kPageHeaderLen) return PGF_ERR_INVALID_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  PGF_TRY(scan_reserve(ctx, s, s.npages + npages));
  const uint32_t len = uint32_t(stride < ctx->page_size ? stride : ctx->page_size);
  const uint64_t first = s.npages;
  const bool pinned = host_ptr_is_pinned(pages);
  auto rollback = [&]() {  // forget the pages admitted by this call
    s.h_descs.resize(first);
    s.rows = first ? s.h_descs.back().row_base + s.h_descs.back().row_count : 0;
    s.npages = first;
  };
  if (pinned) {
    // caller-owned pinned (or registered shared-memory) pages: DMA straight from them, in
    // chunks, so the host-side admission checks of chunk k+1 run under the copy of chunk k
    const uint64_t chunk = 2048;
    for (uint64_t p0 = 0; p0 < npages; p0 += chunk) {
      const uint64_t n = npages - p0 < chunk ? npages - p0 : chunk;
      for (uint64_t p = p0; p < p0 + n; ++p) {
        pgf_status st = scan_admit_page(ctx, s, pages + p * stride, len);
        if (st) {
          cudaStreamSynchronize(ctx->copy_stream);
          rollback();
          return st;
        }
        s.npages++;
      }
      uint8_t* dst = s.d_pages + (first + p0) * uint64_t(ctx->page_size);
      if (stride == ctx->page_size)
        CU(ctx, cudaMemcpyAsync(dst, pages + p0 * stride, n * stride, cudaMemcpyHostToDevice, ctx->copy_stream));
      else
        CU(ctx, cudaMemcpy2DAsync(dst, ctx->page_size, pages + p0 * stride, stride, len, n, cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    s.pending_async += npages;
    return scan_mark_pushed(ctx, s);
  }
  for (uint64_t p = 0; p < npages; ++p) {
    pgf_status st = scan_admit_page(ctx, s, pages + p * stride, len);
    if (st) {
      rollback();
      return st;
    }
    s.npages++;
  }
  uint8_t* dst = s.d_pages + first * uint64_t(ctx->page_size);
  // pageable memory: bounce through the pinned staging chunks (double buffered)
  std::lock_guard<std::mutex> gs(ctx->mu);
  const uint64_t chunk = ctx->staging_pages;
  for (int i = 0; i < 2; ++i)
    if (!ctx->staging[i]) {
      cudaError_t e = cudaMallocHost(&ctx->staging[i], chunk * uint64_t(ctx->page_size));
      if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate pinned staging");
      }
    }
  for (uint64_t p0 = 0; p0 < npages; p0 += chunk) {
    const uint64_t n = npages - p0 < chunk ? npages - p0 : chunk;
    const int b = ctx->staging_next;
    ctx->staging_next ^= 1;
    CU(ctx, cudaEventSynchronize(ctx->staging_ev[b]));
    if (stride == ctx->page_size) {
      std::memcpy(ctx->staging[b], pages + p0 * stride, n * stride);
    } else {
      for (uint64_t p = 0; p < n; ++p)
        std::memcpy(ctx->staging[b] + p * ctx->page_size, pages + (p0 + p) * stride, len);
    }
    CU(ctx, cudaMemcpyAsync(dst + p0 * ctx->page_size, ctx->staging[b], n * uint64_t(ctx->page_size),
                            cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(ctx, cudaEventRecord(ctx->staging_ev[b], ctx->copy_stream));
  }
  return scan_mark_pushed(ctx, s);
}

pgf_status pgf_scan_push_page(pgf_ctx* ctx, uint64_t scan_id, const uint8_t* page, uint32_t len) {
  if (!ctx || !page) return PGF_ERR_INVALID_ARGUMENT;
  return pgf_scan_push_pages(ctx, scan_id, page, 1, len);
}

pgf_status pgf_scan_finish(pgf_ctx* ctx, uint64_t scan_id) {
  NvtxRange nvtx_("pgf:scan_finish");
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  Scan* sp = find_scan(ctx, scan_id);
  if (!sp) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  Scan& s = *sp;
  std::lock_guard<std::mutex> g(s.mu);
  if (s.finished) return PGF_OK;
  // every scan thread ends its own stream (transport_scan_source.rs:251-426): the device-side
  // checks share the context's event and error slot, so finishes of different scans serialise
  // here (lock order: scan, then context -- the same as the staging path of push_pages)
  std::lock_guard<std::mutex> gc(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  PGF_TRY(scan_sync_descs(ctx, s));
  // compute stream waits for the H2D copies of this scan; then the row-level import checks run on device
  if (s.ev_pushed) {
    CU(ctx, cudaStreamWaitEvent(ctx->compute_stream, s.ev_pushed, 0));
  } else {   // (nothing was pushed through the copy stream)
    CU(ctx, cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    CU(ctx, cudaStreamWaitEvent(ctx->compute_stream, ctx->ev_copy, 0));
  }
  PGF_TRY(scan_device_validate(ctx, s));
  s.pending_async = 0;
  s.finished = true;
  return PGF_OK;
}

pgf_status pgf_scan_get_info(pgf_ctx* ctx, uint64_t scan_id, pgf_scan_info* out) {
  if (!ctx || !out) return PGF_ERR_INVALID_ARGUMENT;
  Scan* s = find_scan(ctx, scan_id);
  if (!s) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  out->pages = s->npages;
  out->rows = s->rows;
  out->bytes = s->npages * uint64_t(ctx->page_size);
  out->ncols = uint32_t(s->schema.size());
  out->finished = s->finished;
  return PGF_OK;
}

pgf_status pgf_scan_reset(pgf_ctx* ctx, uint64_t scan_id) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  Scan* s = find_scan(ctx, scan_id);
  if (!s) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  std::lock_guard<std::mutex> g(s->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  // the scan's own copies and every kernel that may still read its pages; copies of OTHER scans stay in flight
  if (s->ev_pushed) CU(ctx, cudaEventSynchronize(s->ev_pushed));
  else CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  s->npages = 0;
  s->rows = 0;
  s->finished = false;
  s->h_descs.clear();
  s->h_classes.clear();
  s->max_page_rows = 0;
  s->descs_dirty = true;
  s->pending_async = 0;
  return PGF_OK;
}

pgf_status pgf_scan_release(pgf_ctx* ctx, uint64_t scan_id) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->scans.find(scan_id);
  if (it == ctx->scans.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->compute_stream);
  if (it->second->ev_pushed) cudaEventDestroy(it->second->ev_pushed);
  if (it->second->d_pages) cudaFree(it->second->d_pages);
  if (it->second->d_descs) cudaFree(it->second->d_descs);
  if (it->second->d_classes) cudaFree(it->second->d_classes);
  ctx->scans.erase(it);
  return PGF_OK;
}

pgf_status pgf_scan_read_pages(pgf_ctx* ctx, uint64_t scan_id, uint64_t first_page, uint64_t npages, uint8_t* out) {
  if (!ctx || (!out && npages)) return PGF_ERR_INVALID_ARGUMENT;
  Scan* s = find_scan(ctx, scan_id);
  if (!s) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  if (first_page + npages > s->npages) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "page range out of bounds");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  CU(ctx, cudaMemcpy(out, s->d_pages + first_page * uint64_t(ctx->page_size), npages * uint64_t(ctx->page_size),
                     cudaMemcpyDeviceToHost));
  return PGF_OK;
}

/* ---- Bloom parameters and lifecycle (host state machine; bits live in HBM) ---- */
pgf_status pgf_bloom_params_new(uint64_t bit_count, uint64_t hash_count, uint64_t seed, pgf_bloom_params* out) {
  if (!out) return PGF_ERR_INVALID_ARGUMENT;
  if (bit_count == 0) return PGF_ERR_BLOOM_ZERO_BIT_COUNT;    // bloom.rs:30-32
  if (hash_count == 0) return PGF_ERR_BLOOM_ZERO_HASH_COUNT;  // bloom.rs:33-35
  if (bit_count > UINT64_MAX - 63) return PGF_ERR_BLOOM_TOO_MANY_BITS;
  out->bit_count = bit_count;
  out->word_count = (bit_count + 63) / 64;
  out->hash_count = hash_count;
  out->seed = seed;
  return PGF_OK;
}

}  // extern "C"

#include <cmath>

extern "C" {

pgf_status pgf_bloom_params_for_expected_items(uint64_t expected_items, double fpr, uint64_t seed,
                                               pgf_bloom_params* out) {
  if (!out) return PGF_ERR_INVALID_ARGUMENT;
  if (expected_items == 0) return PGF_ERR_BLOOM_ZERO_EXPECTED_ITEMS;
  if (!std::isfinite(fpr) || fpr <= 0.0 || fpr >= 1.0) return PGF_ERR_BLOOM_INVALID_FPR;
  const double n = double(expected_items);
  const double ln2 = 0.693147180559945309417232121458176568;
  const double bits = std::ceil(-(n * std::log(fpr)) / (ln2 * ln2));  // bloom.rs:70
  if (bits > 18446744073709551615.0) return PGF_ERR_BLOOM_TOO_MANY_BITS;
  const uint64_t bit_count = uint64_t(bits);
  double k = std::round((double(bit_count) / n) * ln2);  // bloom.rs:75
  if (k < 1.0) k = 1.0;
  return pgf_bloom_params_new(bit_count, uint64_t(k), seed, out);
}

static BloomSlot* find_bloom(pgf_ctx* ctx, uint64_t id) {
  auto it = ctx->blooms.find(id);
  return it == ctx->blooms.end() ? nullptr : &it->second;
}

#define BLOOM_OR_FAIL(ctx, id, var)                                                          \
  if (!(ctx)) return PGF_ERR_INVALID_ARGUMENT;                                               \
  if ((ctx)->sticky) return (ctx)->sticky;                                                   \
  BloomSlot* var = find_bloom((ctx), (id));                                                  \
  if (!var) return (ctx)->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown bloom filter %llu", (unsigned long long)(id));

pgf_status pgf_bloom_create(pgf_ctx* ctx, const pgf_bloom_params* params, uint64_t* bloom_out) {
  if (!ctx || !params || !bloom_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  pgf_bloom_params p{};
  PGF_TRY(pgf_bloom_params_new(params->bit_count, params->hash_count, params->seed, &p));
  if (p.word_count != params->word_count) return ctx->fail(PGF_ERR_BLOOM_INSUFFICIENT_WORDS, "word_count %llu does not match bit_count", (unsigned long long)params->word_count);
  CU(ctx, cudaSetDevice(ctx->device));
  BloomSlot b;
  b.params = p;
  cudaError_t e = cudaMalloc(&b.d_words, p.word_count * 8);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "cannot allocate %llu Bloom words", (unsigned long long)p.word_count);
  }
  CU(ctx, cudaMemsetAsync(b.d_words, 0, p.word_count * 8, ctx->compute_stream));
  PGF_TRY(bloom_make_dev(p, b.d_words, &b.dev));
  const uint64_t id = ctx->next_handle++;
  ctx->blooms[id] = b;
  *bloom_out = id;
  return PGF_OK;
}

pgf_status pgf_bloom_destroy(pgf_ctx* ctx, uint64_t bloom) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->compute_stream);
  cudaFree(b->d_words);
  ctx->blooms.erase(bloom);
  return PGF_OK;
}

pgf_status pgf_bloom_snapshot(pgf_ctx* ctx, uint64_t bloom, uint64_t* generation, int32_t* state) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if (generation) *generation = b->lifecycle >> 2;  // shared.rs:411-416
  if (state) *state = int32_t(b->lifecycle & 3);
  return PGF_OK;
}

pgf_status pgf_bloom_begin_build(pgf_ctx* ctx, uint64_t bloom, uint64_t* generation_out) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  const uint64_t gen = b->lifecycle >> 2;
  const int state = int(b->lifecycle & 3);
  if (state == PGF_RF_BUILDING || state == PGF_RF_READY)  // shared.rs:165-169
    return ctx->fail(PGF_ERR_LIFECYCLE_BUSY, "runtime filter slot is busy: generation %llu state %d", (unsigned long long)gen, state);
  if (gen + 1 > (UINT64_MAX >> 2))                        // shared.rs:171-180
    return ctx->fail(PGF_ERR_LIFECYCLE_GENERATION_EXHAUSTED, "runtime filter generation %llu cannot advance", (unsigned long long)gen);
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemsetAsync(b->d_words, 0, b->params.word_count * 8, ctx->compute_stream));  // bloom.clear()
  b->lifecycle = ((gen + 1) << 2) | PGF_RF_BUILDING;
  if (generation_out) *generation_out = gen + 1;
  ctx->rf_metrics.allocated_total++;
  return PGF_OK;
}

static pgf_status bloom_transition(pgf_ctx* ctx, BloomSlot* b, int from, int to) {
  const uint64_t gen = b->lifecycle >> 2;
  if (int(b->lifecycle & 3) != from)  // shared.rs:377-397 / 244-260
    return ctx->fail(PGF_ERR_LIFECYCLE_INVALID_TRANSITION, "invalid runtime filter transition: expected state %d, observed generation %llu state %d",
                     from, (unsigned long long)gen, int(b->lifecycle & 3));
  b->lifecycle = (gen << 2) | uint64_t(to);
  return PGF_OK;
}

pgf_status pgf_bloom_publish_ready(pgf_ctx* ctx, uint64_t bloom) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  // all inserts were queued on the compute stream; Ready is observable once they are done
  CU(ctx, cudaSetDevice(ctx->device));
  PGF_TRY(bloom_count_bits(ctx, *b));   // synchronises; the fill decides whether fused probes are worth their cost
  PGF_TRY(bloom_transition(ctx, b, PGF_RF_BUILDING, PGF_RF_READY));
  ctx->rf_metrics.ready_total++;
  return PGF_OK;
}
pgf_status pgf_bloom_disable_build(pgf_ctx* ctx, uint64_t bloom) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  return bloom_transition(ctx, b, PGF_RF_BUILDING, PGF_RF_DISABLED);
}
pgf_status pgf_bloom_retire_ready(pgf_ctx* ctx, uint64_t bloom) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  return bloom_transition(ctx, b, PGF_RF_READY, PGF_RF_DISABLED);
}

pgf_status pgf_bloom_insert_keys(pgf_ctx* ctx, uint64_t bloom, const void* keys, int32_t key_width,
                                 const uint8_t* validity, uint64_t n, uint64_t* rows_inserted) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if ((!keys && n) || (key_width != 2 && key_width != 4 && key_width != 8)) return PGF_ERR_INVALID_ARGUMENT;
  if (int(b->lifecycle & 3) != PGF_RF_BUILDING)  // pool.rs:480-494 re-checks Building per key
    return ctx->fail(PGF_ERR_LIFECYCLE_INVALID_TRANSITION, "insert into a filter that is not Building");
  uint64_t ins = 0;
  PGF_TRY(bloom_insert_host_keys(ctx, *b, keys, key_width, validity, n, &ins));
  ctx->rf_metrics.build_rows_total += ins;
  if (rows_inserted) *rows_inserted = ins;
  return PGF_OK;
}

pgf_status pgf_bloom_insert_scan(pgf_ctx* ctx, uint64_t bloom, uint64_t scan_id, uint32_t col, uint64_t* rows_inserted) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if (int(b->lifecycle & 3) != PGF_RF_BUILDING)
    return ctx->fail(PGF_ERR_LIFECYCLE_INVALID_TRANSITION, "insert into a filter that is not Building");
  Scan* s = find_scan(ctx, scan_id);
  if (!s) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  if (!s->finished) return ctx->fail(PGF_ERR_STATE, "scan %llu is not finished", (unsigned long long)scan_id);
  uint64_t ins = 0;
  PGF_TRY(bloom_insert_scan(ctx, *b, *s, col, &ins));
  ctx->rf_metrics.build_rows_total += ins;
  if (rows_inserted) *rows_inserted = ins;
  return PGF_OK;
}

pgf_status pgf_bloom_read_words(pgf_ctx* ctx, uint64_t bloom, uint64_t* words_out, uint64_t nwords) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if (!words_out) return PGF_ERR_INVALID_ARGUMENT;
  if (nwords < b->params.word_count)  // BloomAttachError::InsufficientWords, bloom.rs:168-173
    return ctx->fail(PGF_ERR_BLOOM_INSUFFICIENT_WORDS, "bloom filter storage has %llu words, but %llu are required",
                     (unsigned long long)nwords, (unsigned long long)b->params.word_count);
  CU(ctx, cudaSetDevice(ctx->device));
  // cudaMemcpyDefault: words_out may be a host buffer or (for the multi-GPU OR-merge) a device buffer
  CU(ctx, cudaMemcpyAsync(words_out, b->d_words, b->params.word_count * 8, cudaMemcpyDefault, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  return PGF_OK;
}

pgf_status pgf_bloom_or_words(pgf_ctx* ctx, uint64_t bloom, const uint64_t* words, uint64_t nwords) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if (!words) return PGF_ERR_INVALID_ARGUMENT;
  if (nwords < b->params.word_count)
    return ctx->fail(PGF_ERR_BLOOM_INSUFFICIENT_WORDS, "bloom filter storage has %llu words, but %llu are required",
                     (unsigned long long)nwords, (unsigned long long)b->params.word_count);
  CU(ctx, cudaSetDevice(ctx->device));
  uint64_t* tmp = nullptr;
  CU(ctx, cudaMalloc(&tmp, b->params.word_count * 8));
  cudaError_t e = cudaMemcpyAsync(tmp, words, b->params.word_count * 8, cudaMemcpyHostToDevice, ctx->compute_stream);
  pgf_status st = e == cudaSuccess ? bloom_or_device(ctx, *b, tmp, b->params.word_count, 1)
                                   : ctx->cuda_fail(e, "cudaMemcpyAsync", __FILE__, __LINE__);
  cudaStreamSynchronize(ctx->compute_stream);
  cudaFree(tmp);
  return st;
}

void* pgf_bloom_device_words(pgf_ctx* ctx, uint64_t bloom) {
  if (!ctx) return nullptr;
  BloomSlot* b = find_bloom(ctx, bloom);
  return b ? b->d_words : nullptr;
}

pgf_status pgf_bloom_or_device_words(pgf_ctx* ctx, uint64_t bloom, const void* dev_words, uint64_t nwords,
                                     uint32_t narrays) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if (!dev_words || nwords != b->params.word_count) return PGF_ERR_INVALID_ARGUMENT;
  return bloom_or_device(ctx, *b, dev_words, nwords, narrays);
}

pgf_status pgf_bloom_or_all_reduce(pgf_ctx* ctx, uint64_t bloom) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if (ctx->comm_world == 1) return PGF_OK;
  const uint64_t bytes = b->params.word_count * 8;
  CU(ctx, cudaSetDevice(ctx->device));
  const uint64_t world = uint64_t(ctx->comm_world), rank = uint64_t(ctx->comm_rank);
  if (bytes >= (1u << 20) && b->params.word_count % (2 * world) == 0) {
    // Large filters (the 32 MiB one of SF100's orders): reduce-scatter + all-gather, written with the collectives
    // NCCL has (it has no bitwise-OR reduction).  Rank r owns words [r, r + 1) * chunk: every rank sends every owner
    // its copy of the owner's chunk (grouped send / recv), the owner ORs the `world` copies, and an in-place all-gather
    // hands the finished chunks round.  2 x 7/8 of the array crosses NVLink per rank instead of 7 x the array
    // (0.51 ms for 32 MiB on 8 GPUs with the plain all-gather).
    const uint64_t chunk_words = b->params.word_count / world, chunk_bytes = chunk_words * 8;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ctx->d_xchg_cap < bytes) {
      CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
      if (ctx->d_xchg) cudaFree(ctx->d_xchg);
      ctx->d_xchg = nullptr;
      ctx->d_xchg_cap = 0;
      void* p = nullptr;
      if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange scratch"); }
      ctx->d_xchg = static_cast<uint8_t*>(p);
      ctx->d_xchg_cap = bytes;
    }
    uint64_t soff[64], sbytes[64], roff[64], rbytes[64];
    if (world > 64) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "at most 64 ranks");
    for (uint64_t p = 0; p < world; ++p) {
      soff[p] = p * chunk_bytes; sbytes[p] = p == rank ? 0 : chunk_bytes;     // (the own copy is already in place)
      roff[p] = p * chunk_bytes; rbytes[p] = p == rank ? 0 : chunk_bytes;
    }
    PGF_TRY(comm_all_to_all_v(ctx, b->d_words, soff, sbytes, ctx->d_xchg, roff, rbytes));
    // the copies of MY chunk sit at d_xchg[p * chunk] for p != rank; slot `rank` of the scratch is skipped by pointing
    // the kernel at two runs of arrays
    uint64_t* mine = b->d_words + rank * chunk_words;
    const uint32_t grid = uint32_t(std::min<uint64_t>((chunk_words + 255) / 256, 1184));
    if (rank) PGF_TRY(bloom_or_strided(ctx, mine, reinterpret_cast<const uint64_t*>(ctx->d_xchg), chunk_words, uint32_t(rank), grid));
    if (rank + 1 < world)
      PGF_TRY(bloom_or_strided(ctx, mine, reinterpret_cast<const uint64_t*>(ctx->d_xchg) + (rank + 1) * chunk_words, chunk_words, uint32_t(world - rank - 1), grid));
    return comm_all_gather(ctx, mine, b->d_words, chunk_bytes);   // in place: my chunk is already where it belongs
  }
  // small filters: all-gather of the word arrays + OR of every array into the local one
  if (ctx->d_xchg_cap < bytes * uint64_t(ctx->comm_world)) {
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
    if (ctx->d_xchg) cudaFree(ctx->d_xchg);
    ctx->d_xchg = nullptr;
    ctx->d_xchg_cap = 0;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes * uint64_t(ctx->comm_world)) != cudaSuccess) { cudaGetLastError(); return ctx->fail(PGF_ERR_OUT_OF_MEMORY, "exchange scratch"); }
    ctx->d_xchg = static_cast<uint8_t*>(p);
    ctx->d_xchg_cap = bytes * uint64_t(ctx->comm_world);
  }
  PGF_TRY(comm_all_gather(ctx, b->d_words, ctx->d_xchg, bytes));
  return bloom_or_device(ctx, *b, ctx->d_xchg, b->params.word_count, uint32_t(ctx->comm_world));
}

// RuntimeFilterProbeStats::record (backend_service/src/source.rs:474-493)
static void note_probe(pgf_ctx* ctx, const pgf_probe_stats& st) {
  ctx->rf_metrics.probe_rows_total += st.probe_rows;
  ctx->rf_metrics.probe_rows_rejected_total += st.rejected_rows;
  ctx->rf_metrics.probe_pass_unfiltered_total += st.pass_unfiltered;
}

pgf_status pgf_bloom_probe_keys(pgf_ctx* ctx, uint64_t bloom, uint64_t expected_generation, const void* keys,
                                int32_t key_width, const uint8_t* validity, uint64_t n, uint8_t* decisions_out,
                                pgf_probe_stats* stats) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  if ((!keys && n) || (!decisions_out && n) || (key_width != 2 && key_width != 4 && key_width != 8)) return PGF_ERR_INVALID_ARGUMENT;
  // decision_for_hash: only a Ready filter of the expected generation may reject (shared.rs:350-361)
  const bool ready = (b->lifecycle >> 2) == expected_generation && int(b->lifecycle & 3) == PGF_RF_READY;
  pgf_probe_stats st{};
  PGF_TRY(bloom_probe_host_keys(ctx, *b, ready, keys, key_width, validity, n, decisions_out, &st));
  note_probe(ctx, st);
  if (stats) *stats = st;
  return PGF_OK;
}

pgf_status pgf_bloom_probe_scan(pgf_ctx* ctx, uint64_t bloom, uint64_t expected_generation, uint64_t scan_id,
                                uint32_t col, uint8_t* decisions_out, pgf_probe_stats* stats) {
  BLOOM_OR_FAIL(ctx, bloom, b);
  Scan* s = find_scan(ctx, scan_id);
  if (!s) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown scan %llu", (unsigned long long)scan_id);
  if (!s->finished) return ctx->fail(PGF_ERR_STATE, "scan %llu is not finished", (unsigned long long)scan_id);
  const bool ready = (b->lifecycle >> 2) == expected_generation && int(b->lifecycle & 3) == PGF_RF_READY;
  pgf_probe_stats st{};
  PGF_TRY(bloom_probe_scan(ctx, *b, ready, *s, col, decisions_out, &st));
  note_probe(ctx, st);
  if (stats) *stats = st;
  return PGF_OK;
}

/* ---- pipelines ---- */
pgf_status pgf_pipeline_check(pgf_ctx* ctx, const pgf_pipeline* plan) {
  if (!ctx || !plan) return PGF_ERR_INVALID_ARGUMENT;
  return pipeline_run(ctx, plan, true, nullptr, 0, nullptr, false, nullptr);
}
pgf_status pgf_pipeline_run(pgf_ctx* ctx, const pgf_pipeline* plan, pgf_result** result_out) {
  if (!ctx || !plan || !result_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return pipeline_run(ctx, plan, false, nullptr, 0, nullptr, false, result_out);
}
pgf_status pgf_pipeline_run_partial(pgf_ctx* ctx, const pgf_pipeline* plan, void* dev_state_out,
                                    uint64_t state_capacity_bytes, uint64_t* state_bytes_out, pgf_result** stats_out) {
  if (!ctx || !plan || !dev_state_out || !state_bytes_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return pipeline_run(ctx, plan, false, dev_state_out, state_capacity_bytes, state_bytes_out, true, stats_out);
}
pgf_status pgf_pipeline_merge_partials(pgf_ctx* ctx, const pgf_pipeline* plan, const void* dev_states,
                                       uint64_t state_stride_bytes, uint32_t nstates, pgf_result** result_out) {
  if (!ctx || !plan || !dev_states || !result_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return pipeline_merge(ctx, plan, dev_states, state_stride_bytes, nstates, false, result_out);
}
pgf_status pgf_pipeline_run_partial_async(pgf_ctx* ctx, const pgf_pipeline* plan, void* dev_state_out,
                                          uint64_t state_capacity_bytes) {
  if (!ctx || !plan || !dev_state_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return pipeline_run_partial_async(ctx, plan, dev_state_out, state_capacity_bytes);
}
pgf_status pgf_pipeline_merge_partials_bounded(pgf_ctx* ctx, const pgf_pipeline* plan, const void* dev_states,
                                               uint64_t state_stride_bytes, uint32_t nstates, pgf_result** result_out) {
  if (!ctx || !plan || !dev_states || !result_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return pipeline_merge(ctx, plan, dev_states, state_stride_bytes, nstates, true, result_out);
}
pgf_status pgf_partial_state_bytes(const pgf_pipeline* plan, uint64_t max_groups, uint64_t* bytes_out) {
  if (!plan || !bytes_out) return PGF_ERR_INVALID_ARGUMENT;
  // [count][max_groups x (key words, null mask, accumulators (<= 2 words each), counts)]
  const uint64_t entry = kKeyWords + 1 + uint64_t(plan->nexprs) * 2 + plan->nexprs + 1;
  *bytes_out = (1 + (max_groups ? max_groups : 1) * entry) * 8;
  return PGF_OK;
}
void pgf_result_free(pgf_result* r) {
  if (!r) return;
  delete[] r->keys;
  delete[] r->aggs;
  delete r;
}
/* ---- result pages (ResultPageProducer / BatchPageEncoder) ---- */
pgf_status pgf_result_schema(const pgf_result* r, pgf_column_spec* schema_out, uint32_t* ncols_out) {
  if (!r || !schema_out || !ncols_out) return PGF_ERR_INVALID_ARGUMENT;
  uint32_t n = 0;
  // a group key is nullable iff its source column is (DataFusion: the AggregateExec output field of a
  // group expression keeps the nullability of its input field)
  for (uint32_t k = 0; k < r->nkeys; ++k) schema_out[n++] = pgf_column_spec{uint16_t(r->key_type[k]), uint16_t(r->key_not_null[k] ? 0 : 1)};
  for (uint32_t a = 0; a < r->naggs; ++a) {
    // COUNT never yields NULL; SUM / AVG over no (non-null) rows do, so their columns are nullable
    // whatever this result happens to hold (the receiving schema is fixed by the plan)
    bool is_count;
    if (r->agg_func[a]) {
      is_count = r->agg_func[a] == PGF_AGG_COUNT || r->agg_func[a] == PGF_AGG_COUNT_STAR;
    } else {  // hand-built result without the function: an Int64 column without NULLs is taken for a count
      bool has_null = false;
      for (uint64_t g = 0; g < r->ngroups && !has_null; ++g) has_null = r->aggs[g * r->naggs + a].kind == PGF_V_NULL;
      is_count = r->agg_type[a] == PGF_T_INT64 && !has_null;
    }
    schema_out[n++] = pgf_column_spec{uint16_t(r->agg_type[a]), uint16_t(is_count ? 0 : 1)};
  }
  *ncols_out = n;
  return n ? PGF_OK : PGF_ERR_INVALID_ARGUMENT;
}

namespace {
// one output cell -> its fixed-width slot (integers narrow to the column width; strings become
// inline ByteViews: len + up to 12 bytes, zero padded, page/arrow_layout/src/raw.rs:114-126)
void put_cell(uint8_t* dst, int type, const pgf_value& v) {
  switch (type) {
    case PGF_T_INT16: { const int16_t x = int16_t(v.lo); std::memcpy(dst, &x, 2); break; }
    case PGF_T_INT32: { const int32_t x = int32_t(v.lo); std::memcpy(dst, &x, 4); break; }
    case PGF_T_INT64: std::memcpy(dst, &v.lo, 8); break;
    case PGF_T_FLOAT32: { const float x = float(v.f64); std::memcpy(dst, &x, 4); break; }
    case PGF_T_FLOAT64: std::memcpy(dst, &v.f64, 8); break;
    case PGF_T_DECIMAL128: std::memcpy(dst, &v.lo, 8); std::memcpy(dst + 8, &v.hi, 8); break;
    default: {  // views
      std::memset(dst, 0, 16);
      const int32_t len = v.slen;
      std::memcpy(dst, &len, 4);
      std::memcpy(dst + 4, v.str, size_t(len > 12 ? 12 : len));
    }
  }
}
}  // namespace

pgf_status pgf_result_encode_pages(const pgf_result* r, uint32_t page_size, uint64_t first_row, uint8_t* pages_out,
                                   uint64_t max_pages, uint64_t* npages_out, uint64_t* rows_done) {
  if (!r || !npages_out || (!pages_out && max_pages)) return PGF_ERR_INVALID_ARGUMENT;
  if (page_size <= kPageHeaderLen + 64 || first_row > r->ngroups) return PGF_ERR_INVALID_ARGUMENT;
  pgf_column_spec schema[PGF_MAX_KEYS + PGF_MAX_AGGS];
  uint32_t ncols = 0;
  PGF_TRY(pgf_result_schema(r, schema, &ncols));
  const uint32_t block_size = page_size - kPageHeaderLen;
  uint32_t cap = 0;
  PGF_TRY(fixed_row_cap(schema, ncols, block_size, &cap));
  if (cap == 0) return PGF_ERR_LAYOUT_DOES_NOT_FIT;
  pgf_layout_plan plan;
  PGF_TRY(plan_layout(schema, ncols, cap, block_size, &plan));
  uint64_t row = first_row, pages = 0;
  std::vector<uint8_t> values(size_t(cap) * 16), validity((cap + 7) / 8);
  while (row < r->ngroups && pages < max_pages) {
    const uint32_t n = uint32_t(std::min<uint64_t>(cap, r->ngroups - row));
    uint8_t* page = pages_out + pages * uint64_t(page_size);
    std::memset(page, 0, page_size);
    encode_page_header(uint16_t(PGF_ARROW_LAYOUT_BATCH_KIND), 0, block_size, page);
    uint8_t* block = page + kPageHeaderLen;
    PGF_TRY(init_block(block, block_size, plan));
    for (uint32_t c = 0; c < ncols; ++c) {
      const int type = schema[c].type_tag;
      const uint32_t w = row_width(type);
      std::fill(validity.begin(), validity.end(), uint8_t(0));
      for (uint32_t i = 0; i < n; ++i) {
        const pgf_value& v = c < r->nkeys ? r->keys[(row + i) * r->nkeys + c] : r->aggs[(row + i) * r->naggs + (c - r->nkeys)];
        if (v.kind == PGF_V_NULL) {
          if (!schema[c].nullable) return PGF_ERR_INVALID_ARGUMENT;   // a NULL in a column declared NOT NULL
          std::memset(values.data() + size_t(i) * w, 0, w);
        } else {
          validity[i >> 3] |= uint8_t(1u << (i & 7));
          put_cell(values.data() + size_t(i) * w, type, v);
        }
      }
      PGF_TRY(write_column(block, block_size, c, n, values.data(), schema[c].nullable ? validity.data() : nullptr));
    }
    PGF_TRY(set_row_count(block, block_size, n));
    row += n;
    ++pages;
  }
  *npages_out = pages;
  if (rows_done) *rows_done = row - first_row;
  return PGF_OK;
}

uint32_t pgf_partition_of_key(int64_t key, uint32_t world) { return world ? join_partition(key, world) : 0u; }

pgf_status pgf_join_table_destroy(pgf_ctx* ctx, uint64_t join_table) {
  if (!ctx) return PGF_ERR_INVALID_ARGUMENT;
  auto it = ctx->joins.find(join_table);
  if (it == ctx->joins.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown join table %llu", (unsigned long long)join_table);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->compute_stream);
  ctx->join_free(it->second.d_slots, it->second.alloc_bytes);   // recycled by the next build of a similar size
  ctx->join_free(it->second.d_rows, it->second.rows_alloc_bytes);
  ctx->joins.erase(it);
  return PGF_OK;
}

pgf_status pgf_join_table_get_info(pgf_ctx* ctx, uint64_t join_table, pgf_join_info* out) {
  if (!ctx || !out) return PGF_ERR_INVALID_ARGUMENT;
  auto it = ctx->joins.find(join_table);
  if (it == ctx->joins.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown join table %llu", (unsigned long long)join_table);
  out->rows = it->second.rows;
  out->capacity = it->second.capacity;
  out->row_bytes = it->second.slot_u4 * 16u;
  out->npayload = it->second.npayload;
  return PGF_OK;
}
pgf_status pgf_join_table_export(pgf_ctx* ctx, uint64_t join_table, void* dev_rows_out, uint64_t capacity_rows, uint64_t* rows_out) {
  if (!ctx || !dev_rows_out || !rows_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  auto it = ctx->joins.find(join_table);
  if (it == ctx->joins.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown join table %llu", (unsigned long long)join_table);
  return join_export(ctx, it->second, dev_rows_out, capacity_rows, rows_out);
}
pgf_status pgf_join_table_from_fragments(pgf_ctx* ctx, uint64_t like_table, const void* dev_rows, uint64_t stride_bytes,
                                         const uint64_t* counts, uint32_t nfragments, uint64_t* table_out) {
  if (!ctx || (!dev_rows && nfragments) || (!counts && nfragments) || !table_out) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  auto it = ctx->joins.find(like_table);
  if (it == ctx->joins.end()) return ctx->fail(PGF_ERR_UNKNOWN_HANDLE, "unknown join table %llu", (unsigned long long)like_table);
  return join_from_fragments(ctx, it->second, dev_rows, stride_bytes, counts, nfragments, table_out);
}

pgf_status pgf_gen_scan(pgf_ctx* ctx, uint64_t scan_id, const pgf_gen_spec* spec) {
  if (!ctx || !spec) return PGF_ERR_INVALID_ARGUMENT;
  if (ctx->sticky) return ctx->sticky;
  return gen_scan(ctx, scan_id, spec);
}
pgf_status pgf_gen_schema(int32_t table, pgf_column_spec* schema_out, uint32_t* ncols_out) {
  if (!schema_out || !ncols_out) return PGF_ERR_INVALID_ARGUMENT;
  return gen_schema(table, schema_out, ncols_out);
}

}  // extern "C"
