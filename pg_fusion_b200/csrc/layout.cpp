// Host-side page layout: see layout.hpp.  Reference contract: page/arrow_layout/src/
// {plan,validate,internals,access}.rs, page/import/src/lib.rs, page/transfer/src/page.rs.
#include "layout.hpp"

#include <cstring>
#include <optional>

namespace pgf {
namespace {

// Checked u32 arithmetic in the spirit of the reference's checked_add / SizeOverflow.
struct U32 {
  uint64_t v;
  bool ok() const { return v <= UINT32_MAX; }
};

inline std::optional<uint32_t> round_up(uint64_t v, uint32_t a) {
  const uint64_t r = (v + (a - 1)) & ~uint64_t(a - 1);
  if (r > UINT32_MAX) return std::nullopt;
  return static_cast<uint32_t>(r);
}

// internals.rs:101-114: smallest x >= v with x % a == bias
inline std::optional<uint32_t> round_up_biased(uint64_t v, uint32_t a, uint32_t bias) {
  const uint32_t delta = a - bias;
  auto up = round_up(v + delta, a);
  if (!up || *up < delta) return std::nullopt;
  return *up - delta;
}

// bitmap.rs:4-6 (div_ceil: `(rows + 7) / 8` would wrap for rows > 2^32 - 8 and let a crafted max_rows through)
inline uint32_t bitmap_len(uint32_t rows) { return rows / 8 + (rows % 8 != 0); }

inline std::optional<uint32_t> reserved_values(int type, uint32_t max_rows) {  // types.rs:152-162
  if (type == PGF_T_BOOLEAN) return round_up(bitmap_len(max_rows), kAlign);
  return round_up(uint64_t(max_rows) * row_width(type), kAlign);
}

inline std::optional<uint32_t> front_base_for(uint32_t ncols) {  // plan.rs:37-50
  return round_up_biased(sizeof(BlockHeader) + uint64_t(ncols) * sizeof(ColumnDesc), kAlign, kAlignBias);
}

struct Region {
  uint32_t validity_off, validity_len, values_off, values_len;
};

// Walks the front region exactly like plan.rs:52-79 / validate.rs:109-139.
template <class TypeOf>
pgf_status walk_front(uint32_t ncols, uint32_t max_rows, uint32_t front_base, TypeOf type_of,
                      Region* regions, uint64_t* end) {
  uint64_t cur = front_base;
  for (uint32_t c = 0; c < ncols; ++c) {
    const int t = type_of(c);
    auto vl = round_up(bitmap_len(max_rows), kAlign);
    auto dl = reserved_values(t, max_rows);
    if (!vl || !dl) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
    Region r;
    r.validity_off = static_cast<uint32_t>(cur);
    r.validity_len = *vl;
    cur += *vl;
    if (cur > UINT32_MAX) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
    r.values_off = static_cast<uint32_t>(cur);
    r.values_len = *dl;
    cur += *dl;
    if (cur > UINT32_MAX) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
    if (regions) regions[c] = r;
  }
  *end = cur;
  return PGF_OK;
}

inline BlockHeader load_header(const uint8_t* b) {
  BlockHeader h;
  std::memcpy(&h, b, sizeof h);
  return h;
}
inline ColumnDesc load_desc(const uint8_t* b, uint32_t i) {
  ColumnDesc d;
  std::memcpy(&d, b + sizeof(BlockHeader) + size_t(i) * sizeof(ColumnDesc), sizeof d);
  return d;
}
inline void store_desc(uint8_t* b, uint32_t i, const ColumnDesc& d) {
  std::memcpy(b + sizeof(BlockHeader) + size_t(i) * sizeof(ColumnDesc), &d, sizeof d);
}

inline bool bit(const uint8_t* bm, uint32_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

bool valid_utf8(const uint8_t* s, uint32_t n) {
  for (uint32_t i = 0; i < n;) {
    const uint8_t c = s[i];
    if (c < 0x80) { ++i; continue; }
    uint32_t extra, cp, min;
    if ((c & 0xE0) == 0xC0) { extra = 1; cp = c & 0x1F; min = 0x80; }
    else if ((c & 0xF0) == 0xE0) { extra = 2; cp = c & 0x0F; min = 0x800; }
    else if ((c & 0xF8) == 0xF0) { extra = 3; cp = c & 0x07; min = 0x10000; }
    else return false;
    if (i + extra >= n) return false;
    for (uint32_t k = 1; k <= extra; ++k) {
      if ((s[i + k] & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (s[i + k] & 0x3F);
    }
    if (cp < min || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += extra + 1;
  }
  return true;
}

}  // namespace

pgf_status plan_layout(const pgf_column_spec* specs, uint32_t ncols, uint32_t max_rows,
                       uint32_t block_size, pgf_layout_plan* out) {
  if (ncols > 64) return PGF_ERR_LAYOUT_TOO_MANY_COLUMNS;
  for (uint32_t c = 0; c < ncols; ++c)
    if (!known_type(specs[c].type_tag)) return PGF_ERR_LAYOUT_INVALID_TYPE_TAG;
  auto fb = front_base_for(ncols);
  if (!fb) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
  Region regions[64];
  uint64_t end = 0;
  pgf_status st = walk_front(ncols, max_rows, *fb, [&](uint32_t c) { return int(specs[c].type_tag); },
                             regions, &end);
  if (st) return st;
  if (end > block_size) return PGF_ERR_LAYOUT_DOES_NOT_FIT;
  out->block_size = block_size;
  out->max_rows = max_rows;
  out->front_base = *fb;
  out->pool_base = static_cast<uint32_t>(end);
  out->ncols = ncols;
  for (uint32_t c = 0; c < ncols; ++c) {
    pgf_column_layout& l = out->cols[c];
    l.type_tag = specs[c].type_tag;
    l.flags = uint16_t((specs[c].nullable ? kFlagNullable : 0) | (is_view(specs[c].type_tag) ? kFlagView : 0));
    l.validity_off = regions[c].validity_off;
    l.values_off = regions[c].values_off;
    l.validity_len = regions[c].validity_len;
    l.values_len = regions[c].values_len;
  }
  return PGF_OK;
}

pgf_status fixed_row_cap(const pgf_column_spec* specs, uint32_t ncols, uint32_t block_size,
                         uint32_t* cap) {
  pgf_layout_plan plan;
  pgf_status st = plan_layout(specs, ncols, 0, block_size, &plan);
  if (st) return st;
  if (block_size > UINT32_MAX / 8) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
  uint32_t lo = 0, hi = block_size * 8;  // row_estimator/src/lib.rs:356-359
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo + 1) / 2;
    st = plan_layout(specs, ncols, mid, block_size, &plan);
    if (st == PGF_OK) lo = mid;
    else if (st == PGF_ERR_LAYOUT_DOES_NOT_FIT) hi = mid - 1;
    else return st;
  }
  *cap = lo;
  return PGF_OK;
}

pgf_status validate_block(const uint8_t* block, size_t len, bool allow_ext) {
  if (len < sizeof(BlockHeader)) return PGF_ERR_LAYOUT_BLOCK_SLICE_TOO_SMALL;
  const BlockHeader h = load_header(block);
  if (len < h.block_size) return PGF_ERR_LAYOUT_BLOCK_SLICE_TOO_SMALL;          // validate.rs:90-96
  if (len < sizeof(BlockHeader) + size_t(h.col_count) * sizeof(ColumnDesc))
    return PGF_ERR_LAYOUT_BLOCK_SLICE_TOO_SMALL;                                 // validate.rs:98-104
  if (h.magic != kBlockMagic) return PGF_ERR_LAYOUT_INVALID_MAGIC;              // validate.rs:30-35
  if (h.version != kBlockVersion) return PGF_ERR_LAYOUT_INVALID_VERSION;
  if (h.row_count > h.max_rows) return PGF_ERR_LAYOUT_ROW_COUNT_EXCEEDS_MAX_ROWS;
  auto fb = front_base_for(h.col_count);
  if (!fb) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
  if (h.front_base != *fb) return PGF_ERR_LAYOUT_FRONT_BASE_MISMATCH;
  if (h.front_base > h.pool_base || h.pool_base > h.tail_cursor || h.tail_cursor > h.block_size)
    return PGF_ERR_LAYOUT_INVALID_HEADER_BOUNDS;
  if (h.front_base % kAlign != kAlignBias || h.pool_base % kAlign != kAlignBias)
    return PGF_ERR_LAYOUT_MISALIGNED_FRONT_REGION;
  // descriptors must tile [front_base, pool_base) exactly (validate.rs:141-172)
  uint64_t cur = h.front_base;
  for (uint32_t c = 0; c < h.col_count; ++c) {
    const ColumnDesc d = load_desc(block, c);
    if (!known_type(d.type_tag, allow_ext)) return PGF_ERR_LAYOUT_INVALID_TYPE_TAG;
    if (((d.flags & kFlagView) != 0) != is_view(d.type_tag)) return PGF_ERR_LAYOUT_INCONSISTENT_VIEW_FLAG;
    auto vl = round_up(bitmap_len(h.max_rows), kAlign);
    auto dl = reserved_values(d.type_tag, h.max_rows);
    if (!vl || !dl) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
    if (d.validity_off != cur) return PGF_ERR_LAYOUT_COLUMN_DESC_MISMATCH;
    cur += *vl;
    if (d.values_off != cur) return PGF_ERR_LAYOUT_COLUMN_DESC_MISMATCH;
    cur += *dl;
    if (cur > UINT32_MAX) return PGF_ERR_LAYOUT_SIZE_OVERFLOW;
    if (d.reserved0 != 0) return PGF_ERR_LAYOUT_COLUMN_DESC_MISMATCH;
  }
  if (cur != h.pool_base) return PGF_ERR_LAYOUT_POOL_BASE_MISMATCH;
  return PGF_OK;
}

namespace {

// kind / flags, BlockRef::open, validate_schema: everything import_owned checks before it touches a column
pgf_status check_block_schema(uint16_t kind, uint16_t flags, const uint8_t* block, size_t len,
                              const pgf_column_spec* schema, uint32_t ncols) {
  if (kind != PGF_ARROW_LAYOUT_BATCH_KIND) return PGF_ERR_IMPORT_WRONG_KIND;    // import/src/lib.rs:121-126
  if (flags != 0) return PGF_ERR_IMPORT_UNSUPPORTED_FLAGS;                      // :127-131
  // the Decimal128 extension tag is structurally valid only under a schema that names it (the caller's opt-in);
  // under a reference schema it is InvalidTypeTag, as in TypeTag::from_raw
  bool ext = false;
  for (uint32_t c = 0; c < ncols; ++c) ext |= schema[c].type_tag == PGF_T_DECIMAL128;
  if (pgf_status st = validate_block(block, len, ext)) return st;
  const BlockHeader h = load_header(block);
  if (h.col_count != ncols) return PGF_ERR_IMPORT_SCHEMA_COLUMN_COUNT_MISMATCH; // :209-214
  // validate_schema runs over every column before any column is imported (:134-138, :208-234)
  for (uint32_t c = 0; c < ncols; ++c) {
    const ColumnDesc d = load_desc(block, c);
    if (d.type_tag != schema[c].type_tag) return PGF_ERR_IMPORT_SCHEMA_TYPE_MISMATCH;
    const bool nullable = (d.flags & kFlagNullable) != 0;
    if (nullable != (schema[c].nullable != 0)) return PGF_ERR_IMPORT_SCHEMA_NULLABILITY_MISMATCH;
  }
  return PGF_OK;
}

// import_nulls bounds (:245-262)
pgf_status check_null_count_bounds(const BlockHeader& h, const ColumnDesc& d) {
  const bool nullable = (d.flags & kFlagNullable) != 0;
  if (!nullable ? d.null_count != 0 : d.null_count > h.row_count) return PGF_ERR_IMPORT_INVALID_NULL_COUNT;
  return PGF_OK;
}

}  // namespace

pgf_status check_block_structure(uint16_t kind, uint16_t flags, const uint8_t* block, size_t len,
                                 const pgf_column_spec* schema, uint32_t ncols) {
  if (pgf_status st = check_block_schema(kind, flags, block, len, schema, ncols)) return st;
  const BlockHeader h = load_header(block);
  for (uint32_t c = 0; c < ncols; ++c)
    if (pgf_status st = check_null_count_bounds(h, load_desc(block, c))) return st;
  return PGF_OK;
}

pgf_status check_block_full(uint16_t kind, uint16_t flags, const uint8_t* block, size_t len,
                            const pgf_column_spec* schema, uint32_t ncols) {
  if (pgf_status st = check_block_schema(kind, flags, block, len, schema, ncols)) return st;
  const BlockHeader h = load_header(block);
  const uint32_t pool_capacity = h.block_size - h.pool_base;
  const uint32_t tail_start = h.tail_cursor - h.pool_base;  // access.rs:88-93
  // columns are imported one after the other: the first failing column decides the error (:160-203)
  for (uint32_t c = 0; c < ncols; ++c) {
    const ColumnDesc d = load_desc(block, c);
    if (pgf_status st = check_null_count_bounds(h, d)) return st;
    const bool nullable = (d.flags & kFlagNullable) != 0;
    const uint8_t* validity = block + d.validity_off;
    if (nullable) {  // import/src/lib.rs:264-289
      uint32_t set = 0;
      for (uint32_t r = 0; r < h.row_count; ++r) set += bit(validity, r);
      if (h.row_count - set != d.null_count) return PGF_ERR_IMPORT_NULL_BITMAP_COUNT_MISMATCH;
    }
    if (!is_view(d.type_tag)) continue;
    for (uint32_t r = 0; r < h.row_count; ++r) {  // validate_view_tail :424-452 + arrow view checks
      if (nullable && !bit(validity, r)) continue;
      ByteView v;
      std::memcpy(&v, block + d.values_off + size_t(r) * sizeof(ByteView), sizeof v);
      if (v.len < 0) return PGF_ERR_LAYOUT_NEGATIVE_VIEW_LENGTH;
      const uint32_t n = uint32_t(v.len);
      const uint8_t* bytes = v.data;
      if (n > kViewInline) {
        int32_t index, off;
        std::memcpy(&index, v.data + 4, 4);
        std::memcpy(&off, v.data + 8, 4);
        if (index != 0) return PGF_ERR_LAYOUT_INVALID_VIEW_BUFFER_INDEX;
        if (off < 0) return PGF_ERR_LAYOUT_NEGATIVE_VIEW_OFFSET;
        if (uint64_t(uint32_t(off)) + n > pool_capacity) return PGF_ERR_LAYOUT_VIEW_OFFSET_OUT_OF_BOUNDS;
        if (uint32_t(off) < tail_start) return PGF_ERR_IMPORT_VIEW_OFFSET_BEFORE_ALLOCATED_TAIL;
        bytes = block + h.pool_base + uint32_t(off);
        if (std::memcmp(bytes, v.data, 4) != 0) return PGF_ERR_IMPORT_ARROW_INVALID_VIEW;
      } else {
        for (uint32_t k = n; k < kViewInline; ++k)
          if (v.data[k]) return PGF_ERR_IMPORT_ARROW_INVALID_VIEW;
      }
      if (d.type_tag == PGF_T_UTF8VIEW && !valid_utf8(bytes, n)) return PGF_ERR_IMPORT_ARROW_INVALID_VIEW;
    }
  }
  return PGF_OK;
}

pgf_status init_block(uint8_t* block, size_t len, const pgf_layout_plan& plan) {
  if (len < plan.block_size) return PGF_ERR_LAYOUT_BLOCK_SLICE_TOO_SMALL;
  std::memset(block, 0, plan.block_size);
  BlockHeader h{};
  h.magic = kBlockMagic;
  h.version = kBlockVersion;
  h.block_size = plan.block_size;
  h.max_rows = plan.max_rows;
  h.col_count = uint16_t(plan.ncols);
  h.front_base = plan.front_base;
  h.pool_base = plan.pool_base;
  h.tail_cursor = plan.block_size;  // plan.rs:168-183
  std::memcpy(block, &h, sizeof h);
  for (uint32_t c = 0; c < plan.ncols; ++c) {
    ColumnDesc d{};
    d.type_tag = plan.cols[c].type_tag;
    d.flags = plan.cols[c].flags;
    d.validity_off = plan.cols[c].validity_off;
    d.values_off = plan.cols[c].values_off;
    store_desc(block, c, d);
  }
  return PGF_OK;
}

pgf_status write_column(uint8_t* block, size_t len, uint32_t col, uint32_t nrows, const void* values,
                        const uint8_t* validity) {
  if (pgf_status st = validate_block(block, len)) return st;
  const BlockHeader h = load_header(block);
  if (col >= h.col_count) return PGF_ERR_INVALID_ARGUMENT;
  if (nrows > h.max_rows) return PGF_ERR_LAYOUT_ROW_COUNT_EXCEEDS_MAX_ROWS;
  ColumnDesc d = load_desc(block, col);
  const uint32_t w = row_width(d.type_tag);
  const uint32_t nbytes = w ? nrows * w : bitmap_len(nrows);
  std::memcpy(block + d.values_off, values, nbytes);
  // Writers set validity bits even for non-nullable columns (access.rs:316-323);
  // consumers key off the NULLABLE flag, never the bits.
  uint8_t* vb = block + d.validity_off;
  uint32_t nulls = 0;
  if (validity && (d.flags & kFlagNullable)) {
    std::memcpy(vb, validity, bitmap_len(nrows));
    if (nrows % 8) vb[nrows / 8] &= uint8_t((1u << (nrows % 8)) - 1);
    for (uint32_t r = 0; r < nrows; ++r) nulls += !bit(vb, r);
    // null slots are zeroed like BlockMut::write_null (access.rs:322-337)
    if (nulls && w)
      for (uint32_t r = 0; r < nrows; ++r)
        if (!bit(vb, r)) std::memset(block + d.values_off + size_t(r) * w, 0, w);
  } else {
    std::memset(vb, 0xFF, nrows / 8);
    if (nrows % 8) vb[nrows / 8] = uint8_t((1u << (nrows % 8)) - 1);
  }
  d.null_count = nulls;
  store_desc(block, col, d);
  return PGF_OK;
}

pgf_status set_row_count(uint8_t* block, size_t len, uint32_t nrows) {
  if (pgf_status st = validate_block(block, len)) return st;
  BlockHeader h = load_header(block);
  if (nrows > h.max_rows) return PGF_ERR_LAYOUT_ROW_COUNT_EXCEEDS_MAX_ROWS;
  h.row_count = nrows;
  std::memcpy(block, &h, sizeof h);
  return PGF_OK;
}

// msgpack [magic u32, version u16, kind u16, flags u16, payload_len u32] in rmp's
// fixed-width encodings: 0x95, 0xce+4, 0xcd+2, 0xcd+2, 0xcd+2, 0xce+4 = 20 bytes.
void encode_page_header(uint16_t kind, uint16_t flags, uint32_t payload_len, uint8_t out[20]) {
  auto be32 = [](uint8_t* p, uint32_t v) { p[0] = uint8_t(v >> 24); p[1] = uint8_t(v >> 16); p[2] = uint8_t(v >> 8); p[3] = uint8_t(v); };
  auto be16 = [](uint8_t* p, uint16_t v) { p[0] = uint8_t(v >> 8); p[1] = uint8_t(v); };
  out[0] = 0x95;
  out[1] = 0xce; be32(out + 2, kPageMagic);
  out[6] = 0xcd; be16(out + 7, 1);
  out[9] = 0xcd; be16(out + 10, kind);
  out[12] = 0xcd; be16(out + 13, flags);
  out[15] = 0xce; be32(out + 16, payload_len);
}

pgf_status decode_page_header(const uint8_t in[20], uint16_t* kind, uint16_t* flags,
                              uint32_t* payload_len) {
  auto be32 = [](const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; };
  auto be16 = [](const uint8_t* p) { return uint16_t((p[0] << 8) | p[1]); };
  if (in[0] != 0x95 || in[1] != 0xce || in[6] != 0xcd || in[9] != 0xcd || in[12] != 0xcd || in[15] != 0xce)
    return PGF_ERR_IMPORT_PAGE_HEADER_INVALID;
  if (be32(in + 2) != kPageMagic || be16(in + 7) != 1) return PGF_ERR_IMPORT_PAGE_HEADER_INVALID;
  *kind = be16(in + 10);
  *flags = be16(in + 13);
  *payload_len = be32(in + 16);
  return PGF_OK;
}

}  // namespace pgf
