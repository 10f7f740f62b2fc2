// Host-side page layout of the product library: planner, validator, bulk writer.
// Mirrors pg_fusion's `page/arrow_layout` contract (raw.rs / plan.rs / validate.rs) and
// the `page/transfer` page header; consumed by the scan ingest path and the generator.
#pragma once

#include <cstddef>
#include <cstdint>

#include "../../include/pgf_b200.h"

namespace pgf {

constexpr uint32_t kBlockMagic = 0x32424150u;   // constants.rs:4  "PAB2"
constexpr uint16_t kBlockVersion = 1;           // constants.rs:7
constexpr uint32_t kAlign = 16;                 // constants.rs:10
constexpr uint32_t kAlignBias = 12;             // constants.rs:14
constexpr uint32_t kViewInline = 12;            // constants.rs:17
constexpr uint32_t kPageMagic = 0x50545031u;    // transfer/src/page.rs:8
constexpr uint32_t kPageHeaderLen = 20;         // transfer/src/page.rs:11
constexpr uint16_t kFlagNullable = 1, kFlagView = 2;  // types.rs:44-46

#pragma pack(push, 1)
struct BlockHeader {  // raw.rs:21-46
  uint32_t magic;
  uint16_t version, flags;
  uint32_t block_size, max_rows, row_count;
  uint16_t col_count, reserved0;
  uint32_t front_base, pool_base, tail_cursor, reserved1;
};
struct ColumnDesc {  // raw.rs:69-84
  uint16_t type_tag, flags;
  uint32_t validity_off, values_off, null_count, reserved0;
};
struct ByteView {  // raw.rs:106-110
  int32_t len;
  uint8_t data[12];
};
#pragma pack(pop)
static_assert(sizeof(BlockHeader) == 40 && sizeof(ColumnDesc) == 20 && sizeof(ByteView) == 16,
              "on-page struct sizes are part of the format (arrow_layout/src/tests.rs:7-17)");

inline bool is_view(int t) { return t == PGF_T_UTF8VIEW || t == PGF_T_BINARYVIEW; }
// Reference v1 pages know the tags 1..9 (TypeTag::from_raw, types.rs:93-112); 10 (Decimal128) is this library's
// extension and is accepted only where the caller opted in (a schema / layout plan that names it).
inline bool known_type(int t, bool allow_ext = true) { return t >= PGF_T_BOOLEAN && t <= (allow_ext ? PGF_T_DECIMAL128 : PGF_T_BINARYVIEW); }
// bytes per row of the values buffer; 0 for the bit-packed Boolean (types.rs:139-147)
inline uint32_t row_width(int t) {
  switch (t) {
    case PGF_T_INT16: return 2;
    case PGF_T_INT32: case PGF_T_FLOAT32: return 4;
    case PGF_T_INT64: case PGF_T_FLOAT64: return 8;
    case PGF_T_UUID: case PGF_T_UTF8VIEW: case PGF_T_BINARYVIEW: case PGF_T_DECIMAL128: return 16;
    default: return 0;
  }
}

pgf_status plan_layout(const pgf_column_spec* specs, uint32_t ncols, uint32_t max_rows,
                       uint32_t block_size, pgf_layout_plan* out);
pgf_status fixed_row_cap(const pgf_column_spec* specs, uint32_t ncols, uint32_t block_size,
                         uint32_t* cap);
// Structural validation (header + descriptors tile the front region exactly).  allow_ext = false is the
// reference's BlockRef::open: a descriptor with the Decimal128 extension tag is InvalidTypeTag.
pgf_status validate_block(const uint8_t* block, size_t len, bool allow_ext = true);
// Structural + schema + null_count bounds: everything that must hold before the device may
// touch the page.  Row-level checks (bitmap popcount, views) run on the device.
pgf_status check_block_structure(uint16_t kind, uint16_t flags, const uint8_t* block, size_t len,
                                 const pgf_column_spec* schema, uint32_t ncols);
// The complete import_owned check list on the host.
pgf_status check_block_full(uint16_t kind, uint16_t flags, const uint8_t* block, size_t len,
                            const pgf_column_spec* schema, uint32_t ncols);
pgf_status init_block(uint8_t* block, size_t len, const pgf_layout_plan& plan);
pgf_status write_column(uint8_t* block, size_t len, uint32_t col, uint32_t nrows, const void* values,
                        const uint8_t* validity);
pgf_status set_row_count(uint8_t* block, size_t len, uint32_t nrows);
void encode_page_header(uint16_t kind, uint16_t flags, uint32_t payload_len, uint8_t out[20]);
pgf_status decode_page_header(const uint8_t in[20], uint16_t* kind, uint16_t* flags,
                              uint32_t* payload_len);

}  // namespace pgf
