/*
 * oracle/orc_layout.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see orc.h) for the
 * reference's page format: `page/arrow_layout` (planner, validator, writer),
 * `page/transfer` page header and `page/import` import-time checks.
 *
 * Pinned by the reference's own tests (page/arrow_layout/src/tests.rs,
 * page/import/src/tests.rs) re-expressed in tests/test_oracle_layout.py.
 */
#include "orc.h"

#include <stdlib.h>
#include <string.h>

#define HEADER_SIZE 40u /* size_of::<BlockHeader>, arrow_layout/src/tests.rs:15 */
#define DESC_SIZE 20u   /* size_of::<ColumnDesc>,  arrow_layout/src/tests.rs:12 */
#define VIEW_SIZE 16u   /* size_of::<ByteView>,    arrow_layout/src/tests.rs:9 */

/* page/arrow_layout/src/bitmap.rs:4-6 */
static uint32_t bitmap_bytes(uint32_t rows) { return rows / 8 + (rows % 8 != 0); }

/* page/arrow_layout/src/internals.rs:90-99; returns 0 on overflow via *ok */
static uint32_t align_up(uint32_t v, uint32_t a, int *ok) {
  const uint32_t mask = a - 1;
  if (v > UINT32_MAX - mask) { *ok = 0; return 0; }
  return (v + mask) & ~mask;
}

/* page/arrow_layout/src/internals.rs:101-114 */
static uint32_t align_up_bias(uint32_t v, uint32_t a, uint32_t bias, int *ok) {
  const uint32_t delta = a - bias;
  if (v > UINT32_MAX - delta) { *ok = 0; return 0; }
  const uint32_t up = align_up(v + delta, a, ok);
  if (!*ok || up < delta) { *ok = 0; return 0; }
  return up - delta;
}

static int type_is_view(int t) { return t == ORC_T_UTF8VIEW || t == ORC_T_BINARYVIEW; }

static int type_is_known(int t) { return t >= ORC_T_BOOLEAN && t <= ORC_T_DECIMAL128; }

/* page/arrow_layout/src/types.rs:139-147 (+ the Decimal128 extension: 16 bytes) */
int orc_type_row_width(int t) {
  switch (t) {
    case ORC_T_INT16: return 2;
    case ORC_T_INT32: case ORC_T_FLOAT32: return 4;
    case ORC_T_INT64: case ORC_T_FLOAT64: return 8;
    case ORC_T_UUID: case ORC_T_UTF8VIEW: case ORC_T_BINARYVIEW: case ORC_T_DECIMAL128: return 16;
    default: return 0; /* Boolean: bit-packed */
  }
}

/* page/arrow_layout/src/types.rs:152-162 */
static uint32_t values_reserved_len(int t, uint32_t max_rows, int *ok) {
  if (t == ORC_T_BOOLEAN) return align_up(bitmap_bytes(max_rows), ORC_BUFFER_ALIGNMENT, ok);
  const uint64_t bytes = (uint64_t)max_rows * (uint64_t)orc_type_row_width(t);
  if (bytes > UINT32_MAX) { *ok = 0; return 0; }
  return align_up((uint32_t)bytes, ORC_BUFFER_ALIGNMENT, ok);
}

static uint32_t expected_front_base(uint32_t ncols, int *ok) {
  /* plan.rs:37-50 / validate.rs:49-61 */
  return align_up_bias(HEADER_SIZE + ncols * DESC_SIZE, ORC_BUFFER_ALIGNMENT,
                       ORC_BUFFER_ALIGNMENT_BIAS, ok);
}

/* page/arrow_layout/src/plan.rs:33-93 */
int orc_layout_plan_new(const orc_column_spec *specs, uint32_t ncols, uint32_t max_rows,
                        uint32_t block_size, orc_layout_plan *out) {
  if (ncols > 65535u || ncols > ORC_MAX_COLS) return ORC_LE_TOO_MANY_COLUMNS;
  int ok = 1;
  const uint32_t front_base = expected_front_base(ncols, &ok);
  if (!ok) return ORC_LE_SIZE_OVERFLOW;
  uint64_t cursor = front_base;
  for (uint32_t c = 0; c < ncols; ++c) {
    const int t = specs[c].type_tag;
    if (!type_is_known(t)) return ORC_LE_INVALID_TYPE_TAG;
    uint16_t flags = 0;
    if (specs[c].nullable) flags |= ORC_COLFLAG_NULLABLE; /* types.rs:234-243 */
    if (type_is_view(t)) flags |= ORC_COLFLAG_VIEW;
    /* the validity bitmap is reserved for every column, nullable or not (plan.rs:55-57) */
    const uint32_t validity_len = align_up(bitmap_bytes(max_rows), ORC_BUFFER_ALIGNMENT, &ok);
    const uint32_t values_len = values_reserved_len(t, max_rows, &ok);
    if (!ok) return ORC_LE_SIZE_OVERFLOW;
    orc_column_layout *l = &out->cols[c];
    l->type_tag = (uint16_t)t;
    l->flags = flags;
    l->validity_off = (uint32_t)cursor;
    cursor += validity_len;
    if (cursor > UINT32_MAX) return ORC_LE_SIZE_OVERFLOW;
    l->values_off = (uint32_t)cursor;
    cursor += values_len;
    if (cursor > UINT32_MAX) return ORC_LE_SIZE_OVERFLOW;
    l->validity_len = validity_len;
    l->values_len = values_len;
  }
  if (cursor > block_size) return ORC_LE_LAYOUT_DOES_NOT_FIT;
  out->block_size = block_size;
  out->max_rows = max_rows;
  out->front_base = front_base;
  out->pool_base = (uint32_t)cursor;
  out->ncols = ncols;
  return ORC_OK;
}

/* page/row_estimator/src/lib.rs:353-371: largest max_rows whose plan fits. */
int orc_fixed_row_cap(const orc_column_spec *specs, uint32_t ncols, uint32_t block_size,
                      uint32_t *cap_out) {
  orc_layout_plan plan;
  int rc = orc_layout_plan_new(specs, ncols, 0, block_size, &plan);
  if (rc) return rc;
  if (block_size > UINT32_MAX / 8) return ORC_LE_SIZE_OVERFLOW;
  uint32_t low = 0, high = block_size * 8;
  while (low < high) {
    const uint32_t mid = low + ((high - low) + 1) / 2;
    rc = orc_layout_plan_new(specs, ncols, mid, block_size, &plan);
    if (rc == ORC_OK) low = mid;
    else if (rc == ORC_LE_LAYOUT_DOES_NOT_FIT) high = mid - 1;
    else return rc;
  }
  *cap_out = low;
  return ORC_OK;
}

static void read_header(const uint8_t *block, orc_block_header *h) { memcpy(h, block, HEADER_SIZE); }
static void write_header(uint8_t *block, const orc_block_header *h) { memcpy(block, h, HEADER_SIZE); }
static void read_desc(const uint8_t *block, uint32_t i, orc_column_desc *d) {
  memcpy(d, block + HEADER_SIZE + i * DESC_SIZE, DESC_SIZE); /* internals.rs:38-40 */
}
static void write_desc(uint8_t *block, uint32_t i, const orc_column_desc *d) {
  memcpy(block + HEADER_SIZE + i * DESC_SIZE, d, DESC_SIZE);
}

/* page/arrow_layout/src/access.rs:640-654 + plan.rs:168-183 */
int orc_init_block(uint8_t *block, size_t len, const orc_layout_plan *plan) {
  if (len < plan->block_size) return ORC_LE_BLOCK_SLICE_TOO_SMALL;
  memset(block, 0, plan->block_size);
  orc_block_header h;
  memset(&h, 0, sizeof h);
  h.magic = ORC_BLOCK_MAGIC;
  h.version = ORC_BLOCK_VERSION;
  h.block_size = plan->block_size;
  h.max_rows = plan->max_rows;
  h.row_count = 0;
  h.col_count = (uint16_t)plan->ncols;
  h.front_base = plan->front_base;
  h.pool_base = plan->pool_base;
  h.tail_cursor = plan->block_size;
  write_header(block, &h);
  for (uint32_t c = 0; c < plan->ncols; ++c) {
    orc_column_desc d;
    memset(&d, 0, sizeof d);
    d.type_tag = plan->cols[c].type_tag;
    d.flags = plan->cols[c].flags;
    d.validity_off = plan->cols[c].validity_off;
    d.values_off = plan->cols[c].values_off;
    write_desc(block, c, &d);
  }
  return ORC_OK;
}

/* page/arrow_layout/src/validate.rs:23-83 */
static int validate_header(const orc_block_header *h, uint32_t desc_count) {
  if (h->magic != ORC_BLOCK_MAGIC) return ORC_LE_INVALID_MAGIC;
  if (h->version != ORC_BLOCK_VERSION) return ORC_LE_INVALID_VERSION;
  if (h->row_count > h->max_rows) return ORC_LE_ROW_COUNT_EXCEEDS_MAX_ROWS;
  if (h->col_count != desc_count) return ORC_LE_COLUMN_COUNT_MISMATCH;
  int ok = 1;
  const uint32_t fb = expected_front_base(desc_count, &ok);
  if (!ok) return ORC_LE_SIZE_OVERFLOW;
  if (h->front_base != fb) return ORC_LE_FRONT_BASE_MISMATCH;
  if (h->front_base > h->pool_base || h->pool_base > h->tail_cursor ||
      h->tail_cursor > h->block_size)
    return ORC_LE_INVALID_HEADER_BOUNDS;
  if (h->front_base % ORC_BUFFER_ALIGNMENT != ORC_BUFFER_ALIGNMENT_BIAS ||
      h->pool_base % ORC_BUFFER_ALIGNMENT != ORC_BUFFER_ALIGNMENT_BIAS)
    return ORC_LE_MISALIGNED_FRONT_REGION;
  return ORC_OK;
}

/* internals.rs:11-36 */
static int layout_from_desc(uint32_t max_rows, const orc_column_desc *d, orc_column_layout *l) {
  if (!type_is_known(d->type_tag)) return ORC_LE_INVALID_TYPE_TAG;
  const int is_view_flag = (d->flags & ORC_COLFLAG_VIEW) != 0;
  if (is_view_flag != type_is_view(d->type_tag)) return ORC_LE_INCONSISTENT_VIEW_FLAG;
  int ok = 1;
  l->type_tag = d->type_tag;
  l->flags = d->flags;
  l->validity_off = d->validity_off;
  l->values_off = d->values_off;
  l->validity_len = align_up(bitmap_bytes(max_rows), ORC_BUFFER_ALIGNMENT, &ok);
  l->values_len = values_reserved_len(d->type_tag, max_rows, &ok);
  return ok ? ORC_OK : ORC_LE_SIZE_OVERFLOW;
}

/* BlockRef::open = read header, validate_block_prefix, validate_desc_layout_in_block
 * (access.rs:36-42; validate.rs:85-107,141-172) */
static int validate_impl(const uint8_t *block, size_t len, int allow_ext);
int orc_block_validate(const uint8_t *block, size_t len) { return validate_impl(block, len, 1); }
/* the reference as it is: TypeTag::from_raw knows 1..9 (types.rs:93-112), so the Decimal128 extension tag is
 * InvalidTypeTag */
int orc_block_validate_v1(const uint8_t *block, size_t len) { return validate_impl(block, len, 0); }
static int validate_impl(const uint8_t *block, size_t len, int allow_ext) {
  if (len < HEADER_SIZE) return ORC_LE_BLOCK_SLICE_TOO_SMALL;
  orc_block_header h;
  read_header(block, &h);
  if (len < h.block_size) return ORC_LE_BLOCK_SLICE_TOO_SMALL;
  const size_t prefix = HEADER_SIZE + (size_t)h.col_count * DESC_SIZE;
  if (len < prefix) return ORC_LE_BLOCK_SLICE_TOO_SMALL;
  int rc = validate_header(&h, h.col_count);
  if (rc) return rc;
  uint64_t cursor = h.front_base;
  for (uint32_t c = 0; c < h.col_count; ++c) {
    orc_column_desc d;
    orc_column_layout l;
    read_desc(block, c, &d);
    if (!allow_ext && d.type_tag == ORC_T_DECIMAL128) return ORC_LE_INVALID_TYPE_TAG;
    rc = layout_from_desc(h.max_rows, &d, &l);
    if (rc) return rc;
    if (d.validity_off != cursor) return ORC_LE_COLUMN_DESC_MISMATCH;
    cursor += l.validity_len;
    if (d.values_off != cursor) return ORC_LE_COLUMN_DESC_MISMATCH;
    cursor += l.values_len;
    if (cursor > UINT32_MAX) return ORC_LE_SIZE_OVERFLOW;
    if (d.reserved0 != 0) return ORC_LE_COLUMN_DESC_MISMATCH;
  }
  if (cursor != h.pool_base) return ORC_LE_POOL_BASE_MISMATCH;
  return ORC_OK;
}

/* ---- writer (BlockMut), page/arrow_layout/src/access.rs:236-636 ---- */

static int open_mut(uint8_t *block, size_t len, orc_block_header *h) {
  int rc = orc_block_validate(block, len);
  if (rc) return rc;
  read_header(block, h);
  return ORC_OK;
}

static int col_layout(const uint8_t *block, const orc_block_header *h, uint32_t col,
                      orc_column_layout *l) {
  if (col >= h->col_count) return ORC_LE_COLUMN_INDEX_OUT_OF_BOUNDS;
  orc_column_desc d;
  read_desc(block, col, &d);
  return layout_from_desc(h->max_rows, &d, l);
}

/* bitmap.rs:19-29: silently ignores out-of-range bytes */
static void bitmap_set(uint8_t *bytes, uint32_t nbytes, uint32_t idx, int v) {
  const uint32_t b = idx / 8;
  if (b >= nbytes) return;
  if (v) bytes[b] |= (uint8_t)(1u << (idx % 8));
  else bytes[b] &= (uint8_t)~(1u << (idx % 8));
}

static int bitmap_get(const uint8_t *bytes, uint32_t nbytes, uint32_t idx) {
  const uint32_t b = idx / 8;
  if (b >= nbytes) return 0;
  return (bytes[b] >> (idx % 8)) & 1;
}

int orc_block_set_validity(uint8_t *block, size_t len, uint32_t col, uint32_t row, int valid) {
  orc_block_header h;
  orc_column_layout l;
  int rc = open_mut(block, len, &h);
  if (rc) return rc;
  if ((rc = col_layout(block, &h, col, &l))) return rc;
  bitmap_set(block + l.validity_off, l.validity_len, row, valid);
  return ORC_OK;
}

/* access.rs:316-319,575-599 */
int orc_block_write_fixed(uint8_t *block, size_t len, uint32_t col, uint32_t row,
                          const void *bytes, uint32_t nbytes) {
  orc_block_header h;
  orc_column_layout l;
  int rc = open_mut(block, len, &h);
  if (rc) return rc;
  if ((rc = col_layout(block, &h, col, &l))) return rc;
  const uint32_t w = (uint32_t)orc_type_row_width(l.type_tag);
  if (w == 0) return ORC_LE_INVALID_TYPE_TAG;
  if (nbytes != w) return ORC_LE_INVALID_HEADER_BOUNDS;
  if ((uint64_t)row * w + w > l.values_len) return ORC_LE_INVALID_HEADER_BOUNDS;
  bitmap_set(block + l.validity_off, l.validity_len, row, 1);
  memcpy(block + l.values_off + (size_t)row * w, bytes, w);
  return ORC_OK;
}

/* access.rs:301-313 */
int orc_block_write_bool(uint8_t *block, size_t len, uint32_t col, uint32_t row, int value) {
  orc_block_header h;
  orc_column_layout l;
  int rc = open_mut(block, len, &h);
  if (rc) return rc;
  if ((rc = col_layout(block, &h, col, &l))) return rc;
  if (l.type_tag != ORC_T_BOOLEAN) return ORC_LE_INVALID_TYPE_TAG;
  bitmap_set(block + l.validity_off, l.validity_len, row, 1);
  bitmap_set(block + l.values_off, l.values_len, row, value);
  return ORC_OK;
}

/* access.rs:322-337: clears validity and zeroes the value slot */
int orc_block_write_null(uint8_t *block, size_t len, uint32_t col, uint32_t row) {
  orc_block_header h;
  orc_column_layout l;
  int rc = open_mut(block, len, &h);
  if (rc) return rc;
  if ((rc = col_layout(block, &h, col, &l))) return rc;
  bitmap_set(block + l.validity_off, l.validity_len, row, 0);
  if (l.type_tag == ORC_T_BOOLEAN) {
    bitmap_set(block + l.values_off, l.values_len, row, 0);
  } else {
    const uint32_t w = (uint32_t)orc_type_row_width(l.type_tag);
    if ((uint64_t)row * w + w > l.values_len) return ORC_LE_INVALID_HEADER_BOUNDS;
    memset(block + l.values_off + (size_t)row * w, 0, w);
  }
  return ORC_OK;
}

/* access.rs:341-366 (+ tail_alloc :541-557, ByteView::new_inline/new_outline raw.rs:114-150) */
int orc_block_write_view_bytes(uint8_t *block, size_t len, uint32_t col, uint32_t row,
                               const void *bytes, uint32_t nbytes) {
  orc_block_header h;
  orc_column_layout l;
  int rc = open_mut(block, len, &h);
  if (rc) return rc;
  if ((rc = col_layout(block, &h, col, &l))) return rc;
  if (!type_is_view(l.type_tag)) return ORC_LE_INCONSISTENT_VIEW_FLAG;
  if (row >= h.max_rows) return ORC_LE_ROW_COUNT_EXCEEDS_MAX_ROWS;
  if (nbytes > (uint32_t)INT32_MAX) return ORC_LE_SIZE_OVERFLOW;
  orc_byte_view v;
  memset(&v, 0, sizeof v);
  v.len = (int32_t)nbytes;
  if (nbytes <= ORC_VIEW_INLINE_LEN) {
    memcpy(v.data, bytes, nbytes);
  } else {
    if (h.tail_cursor < nbytes) return ORC_LE_SIZE_OVERFLOW;
    const uint32_t next = h.tail_cursor - nbytes;
    if (next < h.pool_base) return ORC_LE_VIEW_FULL;
    h.tail_cursor = next;
    write_header(block, &h);
    memcpy(block + next, bytes, nbytes);
    const int32_t index = 0; /* SHARED_VIEW_BUFFER_INDEX, constants.rs:22-23 */
    const int32_t off = (int32_t)(next - h.pool_base);
    memcpy(v.data, bytes, 4);
    memcpy(v.data + 4, &index, 4);
    memcpy(v.data + 8, &off, 4);
  }
  bitmap_set(block + l.validity_off, l.validity_len, row, 1);
  memcpy(block + l.values_off + (size_t)row * VIEW_SIZE, &v, VIEW_SIZE);
  return ORC_OK;
}

/* access.rs:443-457 */
int orc_block_commit_current_row(uint8_t *block, size_t len) {
  orc_block_header h;
  int rc = open_mut(block, len, &h);
  if (rc) return rc;
  const uint32_t row = h.row_count;
  if (row >= h.max_rows) return ORC_LE_ROW_COUNT_EXCEEDS_MAX_ROWS;
  for (uint32_t c = 0; c < h.col_count; ++c) {
    orc_column_desc d;
    orc_column_layout l;
    read_desc(block, c, &d);
    if ((rc = layout_from_desc(h.max_rows, &d, &l))) return rc;
    if (!bitmap_get(block + l.validity_off, l.validity_len, row)) {
      d.null_count += 1;
      write_desc(block, c, &d);
    }
  }
  h.row_count = row + 1;
  write_header(block, &h);
  return ORC_OK;
}

/* ---- transfer page header: msgpack [magic u32, version u16, kind u16, flags u16,
 * payload_len u32] with rmp's fixed-width encodings (page/transfer/src/page.rs:20-64):
 * fixarray(5)=0x95, u32=0xce + 4 BE bytes, u16=0xcd + 2 BE bytes; 20 bytes total. ---- */
static void put_be32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static void put_be16(uint8_t *p, uint16_t v) { p[0] = v >> 8; p[1] = (uint8_t)v; }
static uint32_t get_be32(const uint8_t *p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static uint16_t get_be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }

int orc_page_header_encode(uint16_t kind, uint16_t flags, uint32_t payload_len, uint8_t out[20]) {
  out[0] = 0x95;
  out[1] = 0xce; put_be32(out + 2, ORC_PAGE_MAGIC);
  out[6] = 0xcd; put_be16(out + 7, 1); /* PAGE_VERSION */
  out[9] = 0xcd; put_be16(out + 10, kind);
  out[12] = 0xcd; put_be16(out + 13, flags);
  out[15] = 0xce; put_be32(out + 16, payload_len);
  return ORC_OK;
}

/* page/transfer/src/page.rs:66-126 */
int orc_page_header_decode(const uint8_t in[20], uint16_t *kind, uint16_t *flags,
                           uint32_t *payload_len) {
  if (in[0] != 0x95) return ORC_IE_PAGE_HEADER_INVALID;
  if (in[1] != 0xce || get_be32(in + 2) != ORC_PAGE_MAGIC) return ORC_IE_PAGE_HEADER_INVALID;
  if (in[6] != 0xcd || get_be16(in + 7) != 1) return ORC_IE_PAGE_HEADER_INVALID;
  if (in[9] != 0xcd || in[12] != 0xcd || in[15] != 0xce) return ORC_IE_PAGE_HEADER_INVALID;
  *kind = get_be16(in + 10);
  *flags = get_be16(in + 13);
  *payload_len = get_be32(in + 16);
  return ORC_OK;
}

/* ByteView::validate, raw.rs:212-243; len sign raw.rs:153-160; offset sign raw.rs:199-209 */
static int view_validate(const orc_byte_view *v, uint32_t pool_capacity, int64_t *offset_out) {
  *offset_out = -1;
  if (v->len < 0) return ORC_LE_NEGATIVE_VIEW_LENGTH;
  if ((uint32_t)v->len <= ORC_VIEW_INLINE_LEN) return ORC_OK;
  int32_t index, off;
  memcpy(&index, v->data + 4, 4);
  memcpy(&off, v->data + 8, 4);
  if (index != 0) return ORC_LE_INVALID_VIEW_BUFFER_INDEX;
  if (off < 0) return ORC_LE_NEGATIVE_VIEW_OFFSET;
  if ((uint64_t)(uint32_t)off + (uint64_t)(uint32_t)v->len > pool_capacity)
    return ORC_LE_VIEW_OFFSET_OUT_OF_BOUNDS;
  *offset_out = off;
  return ORC_OK;
}

static int utf8_valid(const uint8_t *s, uint32_t n) {
  uint32_t i = 0;
  while (i < n) {
    const uint8_t c = s[i];
    uint32_t need;
    uint32_t cp;
    if (c < 0x80) { ++i; continue; }
    else if ((c & 0xE0) == 0xC0) { need = 1; cp = c & 0x1Fu; }
    else if ((c & 0xF0) == 0xE0) { need = 2; cp = c & 0x0Fu; }
    else if ((c & 0xF8) == 0xF0) { need = 3; cp = c & 0x07u; }
    else return 0;
    if (i + need >= n) return 0;
    for (uint32_t k = 1; k <= need; ++k) {
      if ((s[i + k] & 0xC0) != 0x80) return 0;
      cp = (cp << 6) | (s[i + k] & 0x3Fu);
    }
    if ((need == 1 && cp < 0x80) || (need == 2 && cp < 0x800) || (need == 3 && cp < 0x10000)) return 0;
    if (cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return 0;
    i += need + 1;
  }
  return 1;
}

/* page/import/src/lib.rs:117-206 (+ validate_schema :208-235, import_nulls :237-293,
 * validate_view_tail :424-452) */
int orc_import_check(uint16_t kind, uint16_t flags, const uint8_t *block, size_t len,
                     const orc_column_spec *schema, uint32_t ncols) {
  if (kind != ORC_ARROW_LAYOUT_BATCH_KIND) return ORC_IE_WRONG_KIND;
  if (flags != 0) return ORC_IE_UNSUPPORTED_FLAGS;
  int ext = 0; /* the Decimal128 extension tag exists only under a schema that names it */
  for (uint32_t c = 0; c < ncols; ++c) ext |= schema[c].type_tag == ORC_T_DECIMAL128;
  int rc = validate_impl(block, len, ext);
  if (rc) return rc;
  orc_block_header h;
  read_header(block, &h);
  if (h.col_count != ncols) return ORC_IE_SCHEMA_COLUMN_COUNT_MISMATCH;
  for (uint32_t c = 0; c < ncols; ++c) {
    orc_column_desc d;
    read_desc(block, c, &d);
    if (d.type_tag != schema[c].type_tag) return ORC_IE_SCHEMA_TYPE_MISMATCH;
    if (((d.flags & ORC_COLFLAG_NULLABLE) != 0) != (schema[c].nullable != 0))
      return ORC_IE_SCHEMA_NULLABILITY_MISMATCH;
  }
  const uint32_t allocated_tail_start = h.tail_cursor - h.pool_base; /* access.rs:88-93 */
  const uint32_t pool_capacity = h.block_size - h.pool_base;
  for (uint32_t c = 0; c < ncols; ++c) {
    orc_column_desc d;
    orc_column_layout l;
    read_desc(block, c, &d);
    if ((rc = layout_from_desc(h.max_rows, &d, &l))) return rc;
    const int nullable = (d.flags & ORC_COLFLAG_NULLABLE) != 0;
    if (!nullable) {
      if (d.null_count != 0) return ORC_IE_INVALID_NULL_COUNT;
    } else {
      if (d.null_count > h.row_count) return ORC_IE_INVALID_NULL_COUNT;
      uint32_t set = 0;
      for (uint32_t r = 0; r < h.row_count; ++r)
        set += (uint32_t)bitmap_get(block + l.validity_off, l.validity_len, r);
      if (h.row_count - set != d.null_count) return ORC_IE_NULL_BITMAP_COUNT_MISMATCH;
    }
    if (type_is_view(d.type_tag)) {
      for (uint32_t r = 0; r < h.row_count; ++r) {
        if (nullable && !bitmap_get(block + l.validity_off, l.validity_len, r)) continue;
        orc_byte_view v;
        int64_t off;
        memcpy(&v, block + l.values_off + (size_t)r * VIEW_SIZE, VIEW_SIZE);
        if ((rc = view_validate(&v, pool_capacity, &off))) return rc;
        if (off >= 0 && (uint32_t)off < allocated_tail_start)
          return ORC_IE_VIEW_OFFSET_BEFORE_ALLOCATED_TAIL;
        /* arrow-rs view validation run by {String,Binary}ViewArray::try_new [DF-K] */
        const uint8_t *bytes = v.data;
        if (off >= 0) {
          bytes = block + h.pool_base + (uint32_t)off;
          if (memcmp(bytes, v.data, 4) != 0) return ORC_IE_ARROW_INVALID_VIEW; /* prefix */
        } else {
          for (uint32_t k = (uint32_t)v.len; k < ORC_VIEW_INLINE_LEN; ++k)
            if (v.data[k] != 0) return ORC_IE_ARROW_INVALID_VIEW; /* padding */
        }
        if (d.type_tag == ORC_T_UTF8VIEW && !utf8_valid(bytes, (uint32_t)v.len))
          return ORC_IE_ARROW_INVALID_VIEW;
      }
    }
  }
  return ORC_OK;
}

/* ---- decode pages into contiguous columns (what the operators consume) ---- */

void orc_table_free(orc_table *t) {
  for (uint32_t c = 0; c < t->ncols; ++c) {
    free(t->cols[c].values);
    free(t->cols[c].validity);
    free(t->cols[c].arena);
  }
  memset(t, 0, sizeof *t);
}

int orc_table_from_pages(const uint8_t *pages, uint64_t npages, uint64_t page_stride,
                         const orc_column_spec *schema, uint32_t ncols, orc_table *out) {
  memset(out, 0, sizeof *out);
  if (ncols > ORC_MAX_COLS) return ORC_LE_TOO_MANY_COLUMNS;
  uint64_t total = 0;
  for (uint64_t p = 0; p < npages; ++p) {
    const uint8_t *page = pages + p * page_stride;
    uint16_t kind, flags;
    uint32_t payload_len;
    int rc = orc_page_header_decode(page, &kind, &flags, &payload_len);
    if (rc) return rc;
    if ((uint64_t)payload_len + ORC_PAGE_HEADER_LEN > page_stride) return ORC_LE_BLOCK_SLICE_TOO_SMALL;
    rc = orc_import_check(kind, flags, page + ORC_PAGE_HEADER_LEN, payload_len, schema, ncols);
    if (rc) return rc;
    orc_block_header h;
    read_header(page + ORC_PAGE_HEADER_LEN, &h);
    total += h.row_count;
  }
  out->ncols = ncols;
  out->rows = total;
  for (uint32_t c = 0; c < ncols; ++c) {
    orc_column *col = &out->cols[c];
    const int t = schema[c].type_tag;
    col->type_tag = t;
    col->nullable = schema[c].nullable;
    col->rows = total;
    const size_t w = type_is_view(t) ? sizeof(orc_str) : (size_t)orc_type_row_width(t);
    col->values = calloc(total ? total : 1, w ? w : 1); /* Boolean: byte per row */
    col->validity = NULL;
    col->arena = NULL;
    if (!col->values) return ORC_LE_SIZE_OVERFLOW;
    if (type_is_view(t)) {
      /* size the arena: sum of the value lengths of the valid rows */
      uint64_t bytes = 0;
      for (uint64_t p = 0; p < npages; ++p) {
        const uint8_t *block = pages + p * page_stride + ORC_PAGE_HEADER_LEN;
        orc_block_header h;
        orc_column_desc d;
        orc_column_layout l;
        read_header(block, &h);
        read_desc(block, c, &d);
        layout_from_desc(h.max_rows, &d, &l);
        for (uint32_t r = 0; r < h.row_count; ++r) {
          orc_byte_view v;
          memcpy(&v, block + l.values_off + (size_t)r * VIEW_SIZE, VIEW_SIZE);
          if (v.len > 0) bytes += (uint32_t)v.len;
        }
      }
      col->arena = malloc(bytes ? bytes : 1);
      if (!col->arena) return ORC_LE_SIZE_OVERFLOW;
    }
  }
  uint64_t base = 0;
  uint64_t arena_used[ORC_MAX_COLS];
  memset(arena_used, 0, sizeof arena_used);
  for (uint64_t p = 0; p < npages; ++p) {
    const uint8_t *block = pages + p * page_stride + ORC_PAGE_HEADER_LEN;
    orc_block_header h;
    read_header(block, &h);
    for (uint32_t c = 0; c < ncols; ++c) {
      orc_column *col = &out->cols[c];
      orc_column_desc d;
      orc_column_layout l;
      read_desc(block, c, &d);
      layout_from_desc(h.max_rows, &d, &l);
      const int w = orc_type_row_width(d.type_tag);
      const int nullable_with_nulls = (d.flags & ORC_COLFLAG_NULLABLE) && d.null_count > 0;
      if (type_is_view(d.type_tag)) {
        orc_str *dst = (orc_str *)col->values + base;
        for (uint32_t r = 0; r < h.row_count; ++r) {
          orc_byte_view v;
          memcpy(&v, block + l.values_off + (size_t)r * VIEW_SIZE, VIEW_SIZE);
          const int valid =
              !nullable_with_nulls || bitmap_get(block + l.validity_off, l.validity_len, r);
          const uint32_t n = (valid && v.len > 0) ? (uint32_t)v.len : 0;
          const uint8_t *src = v.data;
          if (n > ORC_VIEW_INLINE_LEN) {
            int32_t off;
            memcpy(&off, v.data + 8, 4);
            src = block + h.pool_base + (uint32_t)off; /* raw.rs:98-104 */
          }
          uint8_t *a = col->arena + arena_used[c];
          memcpy(a, src, n);
          arena_used[c] += n;
          dst[r].len = n;
          dst[r].pad = 0;
          dst[r].ptr = a;
        }
      } else if (w) {
        memcpy((uint8_t *)col->values + base * (uint64_t)w, block + l.values_off,
               (size_t)h.row_count * (size_t)w);
      } else {
        for (uint32_t r = 0; r < h.row_count; ++r)
          ((uint8_t *)col->values)[base + r] =
              (uint8_t)bitmap_get(block + l.values_off, l.values_len, r);
      }
      /* nullability keys off the flag, never the bits (page/import/src/lib.rs:245-254) */
      if (nullable_with_nulls) {
        if (!col->validity) {
          col->validity = malloc(total ? total : 1);
          if (!col->validity) return ORC_LE_SIZE_OVERFLOW;
          memset(col->validity, 1, total);
        }
        for (uint32_t r = 0; r < h.row_count; ++r)
          col->validity[base + r] = (uint8_t)bitmap_get(block + l.validity_off, l.validity_len, r);
      }
    }
    base += h.row_count;
  }
  return ORC_OK;
}

static size_t table_col_width(int t) {
  if (type_is_view(t)) return sizeof(orc_str);
  const int w = orc_type_row_width(t);
  return w ? (size_t)w : 1;
}

/* arrow `take` [DF-K]: gather rows (duplicates allowed). View strings keep pointing
 * into a private copy of the source arena. */
int orc_table_take(const orc_table *in, const uint64_t *rows, uint64_t n, orc_table *out) {
  memset(out, 0, sizeof *out);
  out->ncols = in->ncols;
  out->rows = n;
  for (uint32_t c = 0; c < in->ncols; ++c) {
    const orc_column *s = &in->cols[c];
    orc_column *d = &out->cols[c];
    const size_t w = table_col_width(s->type_tag);
    d->type_tag = s->type_tag;
    d->nullable = s->nullable;
    d->rows = n;
    d->values = calloc(n ? n : 1, w);
    d->validity = s->validity ? malloc(n ? n : 1) : NULL;
    d->arena = NULL;
    if (!d->values) return ORC_LE_SIZE_OVERFLOW;
    if (type_is_view(s->type_tag)) {
      uint64_t bytes = 0;
      for (uint64_t i = 0; i < n; ++i) bytes += ((const orc_str *)s->values)[rows[i]].len;
      d->arena = malloc(bytes ? bytes : 1);
      uint64_t used = 0;
      for (uint64_t i = 0; i < n; ++i) {
        const orc_str *sv = (const orc_str *)s->values + rows[i];
        orc_str *dv = (orc_str *)d->values + i;
        memcpy(d->arena + used, sv->ptr, sv->len);
        dv->len = sv->len;
        dv->pad = 0;
        dv->ptr = d->arena + used;
        used += sv->len;
      }
    } else {
      for (uint64_t i = 0; i < n; ++i)
        memcpy((uint8_t *)d->values + i * w, (const uint8_t *)s->values + rows[i] * w, w);
    }
    if (d->validity)
      for (uint64_t i = 0; i < n; ++i) d->validity[i] = s->validity[rows[i]];
  }
  return ORC_OK;
}

/* arrow `filter_record_batch` [DF-K]: keep rows with keep[i] != 0, order preserved. */
int orc_table_select(const orc_table *in, const uint8_t *keep, orc_table *out) {
  uint64_t n = 0;
  for (uint64_t i = 0; i < in->rows; ++i) n += keep[i] != 0;
  uint64_t *rows = malloc((n ? n : 1) * sizeof *rows);
  if (!rows) return ORC_LE_SIZE_OVERFLOW;
  uint64_t k = 0;
  for (uint64_t i = 0; i < in->rows; ++i)
    if (keep[i]) rows[k++] = i;
  const int rc = orc_table_take(in, rows, n, out);
  free(rows);
  return rc;
}
