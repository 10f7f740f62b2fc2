// Scan ingest, device side: descriptor tables and the row-level import checks of
// ArrowPageDecoder::import_owned (page/import/src/lib.rs:237-293 null-bitmap popcounts,
// :424-452 view tail validation, plus arrow's view validation) executed at HBM speed
// after the pages have landed, one CTA per page.
#include "context.hpp"
#include "layout.hpp"

namespace pgf {

namespace {

struct ValidateParams {
  const uint8_t* pages;
  const PageDesc* descs;
  const LayoutClass* classes;
  uint64_t page_stride;
  uint32_t npages;
  uint32_t ncols;
  uint16_t type_tag[kMaxStageCols];
  uint16_t nullable[kMaxStageCols];
  unsigned long long* first_error;  // packed: (page << 24 | col << 16 | code) minimised
};

__device__ __forceinline__ bool dev_valid_utf8(const uint8_t* s, uint32_t n) {
  for (uint32_t i = 0; i < n;) {
    const uint8_t c = s[i];
    if (c < 0x80) { ++i; continue; }
    uint32_t extra, cp, mn;
    if ((c & 0xE0) == 0xC0) { extra = 1; cp = c & 0x1F; mn = 0x80; }
    else if ((c & 0xF0) == 0xE0) { extra = 2; cp = c & 0x0F; mn = 0x800; }
    else if ((c & 0xF8) == 0xF0) { extra = 3; cp = c & 0x07; mn = 0x10000; }
    else return false;
    if (i + extra >= n) return false;
    for (uint32_t k = 1; k <= extra; ++k) {
      if ((s[i + k] & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (s[i + k] & 0x3F);
    }
    if (cp < mn || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += extra + 1;
  }
  return true;
}

__device__ __forceinline__ void report(const ValidateParams& p, uint32_t page, uint32_t col, uint32_t code) {
  atomicMin(p.first_error, (static_cast<unsigned long long>(page) << 24) | (col << 16) | code);
}

__global__ void __launch_bounds__(256) validate_pages_kernel(const ValidateParams p) {
  __shared__ uint32_t s_set;
  for (uint32_t page = blockIdx.x; page < p.npages; page += gridDim.x) {
    const uint8_t* block = p.pages + page * p.page_stride + kPageHeaderLen;
    const BlockHeader* h = reinterpret_cast<const BlockHeader*>(block);
    // Everything that indexes memory below comes from the descriptors the HOST built when it admitted the page
    // (row count, column offsets, pool base), not from the copy of the header that landed in HBM: the copy is
    // asynchronous, and a page that changed between admission and DMA must not steer the kernel out of bounds.
    // A header that no longer agrees with its descriptor is reported and the page skipped.
    const PageDesc pd = p.descs[page];
    const LayoutClass& lc = p.classes[pd.layout_class];
    bool same = h->row_count == pd.row_count && h->pool_base + kPageHeaderLen == lc.pool_base && h->col_count == p.ncols &&
                uint64_t(h->block_size) + kPageHeaderLen <= p.page_stride && h->tail_cursor >= h->pool_base && h->tail_cursor <= h->block_size;
    for (uint32_t c = 0; same && c < p.ncols; ++c) {
      const ColumnDesc* d = reinterpret_cast<const ColumnDesc*>(block + sizeof(BlockHeader) + c * sizeof(ColumnDesc));
      same = d->values_off + kPageHeaderLen == lc.values_off[c] && d->validity_off + kPageHeaderLen == lc.validity_off[c];
    }
    if (!same) {
      if (threadIdx.x == 0) report(p, page, 0, PGF_ERR_IMPORT_PAGE_HEADER_INVALID);
      continue;
    }
    const uint32_t rows = pd.row_count;
    const uint32_t pool_capacity = h->block_size - h->pool_base;
    const uint32_t tail_start = h->tail_cursor - h->pool_base;
    for (uint32_t c = 0; c < p.ncols; ++c) {
      const ColumnDesc* d = reinterpret_cast<const ColumnDesc*>(block + sizeof(BlockHeader) + c * sizeof(ColumnDesc));
      const uint8_t* validity = block + d->validity_off;
      if (p.nullable[c]) {
        if (threadIdx.x == 0) s_set = 0;
        __syncthreads();
        uint32_t set = 0;
        const uint32_t nbytes = (rows + 7) / 8;
        for (uint32_t b = threadIdx.x; b < nbytes; b += blockDim.x) {
          uint32_t v = validity[b];
          if (b == nbytes - 1 && (rows & 7)) v &= (1u << (rows & 7)) - 1;
          set += __popc(v);
        }
        for (int o = 16; o; o >>= 1) set += __shfl_xor_sync(0xffffffffu, set, o);
        if ((threadIdx.x & 31) == 0 && set) atomicAdd(&s_set, set);
        __syncthreads();
        if (threadIdx.x == 0 && rows - s_set != d->null_count)
          report(p, page, c, PGF_ERR_IMPORT_NULL_BITMAP_COUNT_MISMATCH);
        __syncthreads();
      }
      const int t = p.type_tag[c];
      if (t != PGF_T_UTF8VIEW && t != PGF_T_BINARYVIEW) continue;
      const bool has_nulls = p.nullable[c] && d->null_count;
      for (uint32_t r = threadIdx.x; r < rows; r += blockDim.x) {
        if (has_nulls && !((validity[r >> 3] >> (r & 7)) & 1)) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(block + d->values_off + size_t(r) * 16);
        const int32_t len = int32_t(raw.x);
        if (len < 0) { report(p, page, c, PGF_ERR_LAYOUT_NEGATIVE_VIEW_LENGTH); continue; }
        uint8_t inl[12];
        memcpy(inl, &raw.y, 12);
        const uint8_t* bytes = inl;
        if (uint32_t(len) > kViewInline) {
          const int32_t index = int32_t(raw.z), off = int32_t(raw.w);
          if (index != 0) { report(p, page, c, PGF_ERR_LAYOUT_INVALID_VIEW_BUFFER_INDEX); continue; }
          if (off < 0) { report(p, page, c, PGF_ERR_LAYOUT_NEGATIVE_VIEW_OFFSET); continue; }
          if (uint64_t(uint32_t(off)) + uint32_t(len) > pool_capacity) { report(p, page, c, PGF_ERR_LAYOUT_VIEW_OFFSET_OUT_OF_BOUNDS); continue; }
          if (uint32_t(off) < tail_start) { report(p, page, c, PGF_ERR_IMPORT_VIEW_OFFSET_BEFORE_ALLOCATED_TAIL); continue; }
          bytes = block + h->pool_base + uint32_t(off);
          bool same = true;
          for (int k = 0; k < 4; ++k) same &= bytes[k] == inl[k];
          if (!same) { report(p, page, c, PGF_ERR_IMPORT_ARROW_INVALID_VIEW); continue; }
        } else {
          bool pad_ok = true;
          for (uint32_t k = uint32_t(len); k < kViewInline; ++k) pad_ok &= inl[k] == 0;
          if (!pad_ok) { report(p, page, c, PGF_ERR_IMPORT_ARROW_INVALID_VIEW); continue; }
        }
        if (t == PGF_T_UTF8VIEW && !dev_valid_utf8(bytes, uint32_t(len))) report(p, page, c, PGF_ERR_IMPORT_ARROW_INVALID_VIEW);
      }
    }
  }
}

}  // namespace

// Upload the per-page descriptors / layout classes built by the host at admission.
pgf_status scan_sync_descs(pgf_ctx* ctx, Scan& s) {
  if (!s.descs_dirty) return PGF_OK;
  if (s.h_descs.size() > s.d_descs_cap) {
    if (s.d_descs) CU(ctx, cudaFree(s.d_descs));
    s.d_descs_cap = s.h_descs.size() * 2;
    CU(ctx, cudaMalloc(&s.d_descs, s.d_descs_cap * sizeof(PageDesc)));
  }
  if (s.h_classes.size() > s.d_classes_cap) {
    if (s.d_classes) CU(ctx, cudaFree(s.d_classes));
    s.d_classes_cap = s.h_classes.size() * 2;
    CU(ctx, cudaMalloc(&s.d_classes, s.d_classes_cap * sizeof(LayoutClass)));
  }
  if (!s.h_descs.empty()) {
    // (on the compute stream: behind the page copies of every scan on the copy stream this small upload would wait
    // for all of them)
    CU(ctx, cudaMemcpyAsync(s.d_descs, s.h_descs.data(), s.h_descs.size() * sizeof(PageDesc), cudaMemcpyHostToDevice, ctx->compute_stream));
    CU(ctx, cudaMemcpyAsync(s.d_classes, s.h_classes.data(), s.h_classes.size() * sizeof(LayoutClass), cudaMemcpyHostToDevice, ctx->compute_stream));
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));  // h_descs is pageable
  }
  s.descs_dirty = false;
  return PGF_OK;
}

pgf_status scan_device_validate(pgf_ctx* ctx, Scan& s) {
  bool needed = false;
  for (auto& c : s.schema) needed |= c.nullable || is_view(c.type_tag);
  if (!needed || s.npages == 0) {
    CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
    return PGF_OK;
  }
  ValidateParams p{};
  p.pages = s.d_pages;
  p.descs = s.d_descs;
  p.classes = s.d_classes;
  p.page_stride = ctx->page_size;
  p.npages = uint32_t(s.npages);
  p.ncols = uint32_t(s.schema.size());
  for (uint32_t c = 0; c < p.ncols; ++c) {
    p.type_tag[c] = s.schema[c].type_tag;
    p.nullable[c] = s.schema[c].nullable;
  }
  unsigned long long* d_err = reinterpret_cast<unsigned long long*>(ctx->d_flags + 16);
  unsigned long long* h_err = reinterpret_cast<unsigned long long*>(ctx->h_flags + 16);
  p.first_error = d_err;
  CU(ctx, cudaMemsetAsync(d_err, 0xFF, 8, ctx->compute_stream));
  const uint32_t grid = uint32_t(s.npages < uint64_t(ctx->sm_count) * 8 ? s.npages : uint64_t(ctx->sm_count) * 8);
  validate_pages_kernel<<<grid, 256, 0, ctx->compute_stream>>>(p);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(h_err, d_err, 8, cudaMemcpyDeviceToHost, ctx->compute_stream));
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  if (*h_err != ~0ull) {
    const unsigned long long e = *h_err;
    const pgf_status code = pgf_status(e & 0xFFFF);
    return ctx->fail(code, "scan %llu page %llu column %llu rejected by import checks (status %d)",
                     (unsigned long long)s.id, e >> 24, (e >> 16) & 0xFF, int(code));
  }
  return PGF_OK;
}

}  // namespace pgf
