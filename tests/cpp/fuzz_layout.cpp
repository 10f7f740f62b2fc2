// Differential fuzzing of the page validators.  Pages come from other processes' shared memory, so
// the admission checks (pg_fusion_b200/csrc/layout.cpp: validate_block, check_block_structure,
// check_block_full -- BlockRef::open and ArrowPageDecoder::import_owned of the reference) must
// (a) never read outside the slice they are given and (b) give the verdict the reference gives.
//
// The harness links the product's layout.cpp and the oracle's orc_layout.c (test infrastructure;
// the checker here, pinned by the reference's own layout / import tests), both built with
// -fsanitize=address,undefined.  Valid blocks of several schemas are mutated (header fields,
// descriptors, view slots, bitmaps, truncation) and handed to both in exactly-sized heap buffers:
// any disagreement in the status code, and any sanitizer report, fails.
//   fuzz_layout <iterations> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../pg_fusion_b200/csrc/layout.hpp"
extern "C" {
#include "../../oracle/orc.h"
}

namespace {

struct Schema {
  std::vector<pgf_column_spec> cols;
};

std::mt19937_64 rng;
uint64_t pick(uint64_t n) { return n ? rng() % n : 0; }

// One valid block: `rows` rows of pseudo-random values, inline views only (what the writer emits).
std::vector<uint8_t> make_block(const Schema& s, uint32_t block_size, uint32_t max_rows, uint32_t rows) {
  pgf_layout_plan plan;
  if (pgf::plan_layout(s.cols.data(), uint32_t(s.cols.size()), max_rows, block_size, &plan) != PGF_OK) return {};
  std::vector<uint8_t> block(block_size, 0);
  if (pgf::init_block(block.data(), block.size(), plan) != PGF_OK) return {};
  for (uint32_t c = 0; c < s.cols.size(); ++c) {
    const int type = s.cols[c].type_tag;
    const uint32_t w = pgf::row_width(type);
    std::vector<uint8_t> values(size_t(rows) * (w ? w : 1) + 16, 0), validity((rows + 7) / 8 + 1, 0);
    for (uint32_t r = 0; r < rows; ++r) {
      const bool valid = !s.cols[c].nullable || pick(4) != 0;
      if (valid) validity[r >> 3] |= uint8_t(1u << (r & 7));
      if (!w) {  // Boolean: bit-packed values
        if (pick(2)) values[r >> 3] |= uint8_t(1u << (r & 7));
        continue;
      }
      uint8_t* v = values.data() + size_t(r) * w;
      if (pgf::is_view(type)) {
        const int32_t len = valid ? int32_t(pick(13)) : 0;
        std::memcpy(v, &len, 4);
        for (int32_t i = 0; i < len; ++i) v[4 + i] = uint8_t('a' + pick(26));
      } else if (valid) {
        for (uint32_t i = 0; i < w; ++i) v[i] = uint8_t(rng());
        if (type == PGF_T_FLOAT32 || type == PGF_T_FLOAT64) v[w - 1] &= 0x3F;  // finite
      }
    }
    if (pgf::write_column(block.data(), block.size(), c, rows, values.data(), s.cols[c].nullable ? validity.data() : nullptr) != PGF_OK) return {};
  }
  if (pgf::set_row_count(block.data(), block.size(), rows) != PGF_OK) return {};
  return block;
}

// The same kind of block through the oracle's row-at-a-time writer (access.rs:316-457), which also
// emits out-of-line views: strings longer than 12 bytes live in the tail arena that grows down from
// block_size, the slot carries {len, 4-byte prefix, buffer 0, offset from pool_base}.
std::vector<uint8_t> make_block_rowwise(const Schema& s, uint32_t block_size, uint32_t max_rows, uint32_t rows) {
  std::vector<orc_column_spec> os;
  for (const auto& c : s.cols) os.push_back({c.type_tag, c.nullable});
  orc_layout_plan plan;
  if (orc_layout_plan_new(os.data(), uint32_t(os.size()), max_rows, block_size, &plan) != 0) return {};
  std::vector<uint8_t> block(block_size, 0);
  if (orc_init_block(block.data(), block.size(), &plan) != 0) return {};
  for (uint32_t r = 0; r < rows; ++r) {
    for (uint32_t c = 0; c < s.cols.size(); ++c) {
      const int type = s.cols[c].type_tag;
      if (s.cols[c].nullable && pick(4) == 0) {
        if (orc_block_write_null(block.data(), block.size(), c, r) != 0) return {};
        continue;
      }
      int rc = 0;
      if (type == PGF_T_BOOLEAN) {
        rc = orc_block_write_bool(block.data(), block.size(), c, r, int(pick(2)));
      } else if (pgf::is_view(type)) {
        uint8_t text[64];
        const uint32_t n = uint32_t(pick(3) ? pick(13) : 13 + pick(40));
        for (uint32_t i = 0; i < n; ++i) text[i] = uint8_t('a' + pick(26));
        rc = orc_block_write_view_bytes(block.data(), block.size(), c, r, text, n);
        if (rc != 0) rc = orc_block_write_view_bytes(block.data(), block.size(), c, r, text, 3);  // arena full: a short one
      } else {
        uint8_t v[16];
        const uint32_t w = pgf::row_width(type);
        for (uint32_t i = 0; i < w; ++i) v[i] = uint8_t(rng());
        if (type == PGF_T_FLOAT32 || type == PGF_T_FLOAT64) v[w - 1] &= 0x3F;
        rc = orc_block_write_fixed(block.data(), block.size(), c, r, v, w);
      }
      if (rc != 0) return {};
    }
    if (orc_block_commit_current_row(block.data(), block.size()) != 0) return {};
  }
  return block;
}

uint32_t interesting(uint32_t original, uint32_t len) {
  switch (pick(12)) {
    case 0: return 0;
    case 1: return 1;
    case 2: return 0xFFFFFFFFu;
    case 3: return 0x7FFFFFFFu;
    case 4: return 0x80000000u;
    case 5: return len;
    case 6: return len - 1;
    case 7: return len + 1;
    case 8: return original + 1;
    case 9: return original - 1;
    case 10: return original + 16;
    default: return uint32_t(rng());
  }
}

void mutate(std::vector<uint8_t>& b, const Schema& s) {
  const uint32_t len = uint32_t(b.size());
  auto put32 = [&](size_t off, uint32_t v) { if (off + 4 <= b.size()) std::memcpy(b.data() + off, &v, 4); };
  auto get32 = [&](size_t off) { uint32_t v = 0; if (off + 4 <= b.size()) std::memcpy(&v, b.data() + off, 4); return v; };
  auto put16 = [&](size_t off, uint16_t v) { if (off + 2 <= b.size()) std::memcpy(b.data() + off, &v, 2); };
  switch (pick(8)) {
    case 0: {  // a 32-bit header field: magic, block_size, max_rows, row_count, front_base, pool_base, tail_cursor, reserved1
      static const size_t offs[] = {0, 8, 12, 16, 24, 28, 32, 36};
      const size_t o = offs[pick(8)];
      put32(o, interesting(get32(o), len));
      break;
    }
    case 1: {  // a 16-bit header field: version, flags, col_count, reserved0
      static const size_t offs[] = {4, 6, 20, 22};
      put16(offs[pick(4)], uint16_t(pick(2) ? pick(70) : rng()));
      break;
    }
    case 2: {  // a descriptor field
      const size_t d = 40 + 20 * pick(s.cols.size() + 1);
      if (pick(3) == 0) put16(d + 2 * pick(2), uint16_t(pick(2) ? pick(12) : rng()));
      else { const size_t o = d + 4 + 4 * pick(4); put32(o, interesting(get32(o), len)); }
      break;
    }
    case 3:  // any byte
      if (!b.empty()) b[pick(b.size())] ^= uint8_t(1u << pick(8));
      break;
    case 4:  // a byte in the front region (bitmaps / values / view slots live there)
      if (b.size() > 64) b[40 + pick(std::min<size_t>(b.size() - 40, 4096))] = uint8_t(rng());
      break;
    case 5: {  // a view slot becomes an out-of-line reference with arbitrary length / buffer / offset
      const uint32_t pool_base = get32(28);
      for (size_t c = 0; c < s.cols.size(); ++c) {
        if (!pgf::is_view(s.cols[c].type_tag) || pick(2)) continue;
        const uint32_t values_off = get32(40 + 20 * c + 8);
        const size_t slot = size_t(values_off) + 16 * pick(4);
        put32(slot, pick(2) ? uint32_t(13 + pick(200)) : interesting(0, len));
        put32(slot + 8, pick(4) ? 0u : uint32_t(rng()));
        put32(slot + 12, pick(2) ? uint32_t(pick(len)) : interesting(pool_base, len));
        break;
      }
      break;
    }
    case 6:  // truncate the slice
      b.resize(pick(4) ? pick(b.size() + 1) : pick(120));
      break;
    default:  // move tail_cursor into the block so that out-of-line views may look allocated
      put32(32, uint32_t(get32(28) + pick(len)));
      break;
  }
}

}  // namespace

int main(int argc, char** argv) {
  const uint64_t iterations = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 20000;
  rng.seed(argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1);
  const std::vector<Schema> schemas = {
      {{{PGF_T_INT64, 0}}},
      {{{PGF_T_FLOAT64, 0}, {PGF_T_FLOAT64, 1}, {PGF_T_UTF8VIEW, 0}}},
      {{{PGF_T_BOOLEAN, 1}, {PGF_T_INT16, 0}, {PGF_T_INT32, 1}, {PGF_T_INT64, 1}, {PGF_T_FLOAT32, 0}, {PGF_T_FLOAT64, 1},
        {PGF_T_UUID, 0}, {PGF_T_UTF8VIEW, 1}, {PGF_T_BINARYVIEW, 0}}},
      {{{PGF_T_DECIMAL128, 1}, {PGF_T_INT32, 0}, {PGF_T_UTF8VIEW, 1}}},
      {{{PGF_T_BOOLEAN, 1}}},                       // bitmaps only: the one layout where a huge max_rows does not overflow the values
      {{{PGF_T_BOOLEAN, 0}, {PGF_T_BOOLEAN, 1}}},
      {{}},
  };
  uint64_t disagreements = 0, accepted = 0, rejected = 0;
  uint64_t histogram[256] = {0};
  uint64_t rowwise_blocks = 0;
  for (uint64_t it = 0; it < iterations; ++it) {
    const Schema& s = schemas[pick(schemas.size())];
    const uint32_t block_size = uint32_t(1024 + 16 * pick(200));
    uint32_t cap = 0;
    if (pgf::fixed_row_cap(s.cols.data(), uint32_t(s.cols.size()), block_size, &cap) != PGF_OK) continue;
    const uint32_t max_rows = s.cols.empty() ? uint32_t(pick(100)) : uint32_t(1 + pick(cap ? cap : 1));
    if (!s.cols.empty() && cap == 0) continue;
    const uint32_t rows = uint32_t(pick(max_rows + 1));
    const bool rowwise = pick(3) == 0;
    std::vector<uint8_t> valid = rowwise ? make_block_rowwise(s, block_size, max_rows, rows) : make_block(s, block_size, max_rows, rows);
    if (rowwise && !valid.empty()) ++rowwise_blocks;
    if (valid.empty()) continue;
    std::vector<uint8_t> m = valid;
    const uint64_t nmut = pick(4);  // 0 = the valid block itself
    for (uint64_t k = 0; k < nmut; ++k) mutate(m, s);
    // exactly-sized heap copies: an over-read is a sanitizer report, not luck
    uint8_t* exact = static_cast<uint8_t*>(std::malloc(m.size() ? m.size() : 1));
    if (!m.empty()) std::memcpy(exact, m.data(), m.size());
    uint16_t kind = PGF_ARROW_LAYOUT_BATCH_KIND, flags = 0;
    if (pick(50) == 0) kind = uint16_t(rng());
    if (pick(50) == 0) flags = uint16_t(1 + pick(3));
    std::vector<orc_column_spec> oschema;
    std::vector<pgf_column_spec> pschema = s.cols;
    if (!pschema.empty() && pick(30) == 0) pschema[pick(pschema.size())].nullable ^= 1;                    // schema nullability mismatch
    if (!pschema.empty() && pick(30) == 0) pschema[pick(pschema.size())].type_tag = uint16_t(1 + pick(10)); // schema type mismatch
    if (pick(40) == 0) pschema.push_back({PGF_T_INT32, 0});                                                 // column count mismatch
    for (const auto& c : pschema) oschema.push_back({c.type_tag, c.nullable});
    const int v_prod = pgf::validate_block(exact, m.size());
    const int v_orc = orc_block_validate(exact, m.size());
    const int f_prod = pgf::check_block_full(kind, flags, exact, m.size(), pschema.data(), uint32_t(pschema.size()));
    const int f_orc = orc_import_check(kind, flags, exact, m.size(), oschema.data(), uint32_t(oschema.size()));
    // the admission subset (what the host checks before the copy; row-level checks run on the device)
    // never rejects what the full list accepts
    const int s_prod = pgf::check_block_structure(kind, flags, exact, m.size(), pschema.data(), uint32_t(pschema.size()));
    const bool subset_ok = f_prod != PGF_OK || s_prod == PGF_OK;
    if (v_prod != v_orc || f_prod != f_orc || !subset_ok) {
      if (disagreements < 10)
        std::printf("DISAGREE it=%llu len=%zu ncols=%zu validate %d/%d import %d/%d structure %d\n", (unsigned long long)it, m.size(),
                    s.cols.size(), v_prod, v_orc, f_prod, f_orc, s_prod);
      ++disagreements;
    }
    (f_prod == PGF_OK ? accepted : rejected)++;
    if (f_prod >= 0 && f_prod < 256) ++histogram[f_prod];
    if (nmut == 0 && kind == PGF_ARROW_LAYOUT_BATCH_KIND && flags == 0 && pschema.size() == s.cols.size() &&
        (s.cols.empty() || std::memcmp(pschema.data(), s.cols.data(), s.cols.size() * sizeof(pgf_column_spec)) == 0) && f_prod != PGF_OK) {
      std::printf("VALID BLOCK REJECTED it=%llu status %d\n", (unsigned long long)it, f_prod);
      ++disagreements;
    }
    std::free(exact);
  }
  // ---- planner: LayoutPlan::new and the fixed row cap for random schemas, row counts (up to overflow) and block sizes
  uint64_t plans = 0;
  for (uint64_t it = 0; it < iterations; ++it) {
    const uint32_t ncols = uint32_t(pick(3) ? pick(12) : pick(70));
    std::vector<pgf_column_spec> ps(ncols);
    std::vector<orc_column_spec> os(ncols);
    for (uint32_t c = 0; c < ncols; ++c) {
      // (a ColumnSpec of the reference holds a TypeTag enum: unknown tags are rejected when the spec is made,
      // before any planning -- so they are exercised alone, without a competing overflow, further down)
      ps[c] = {uint16_t(1 + pick(10)), uint16_t(pick(2))};
      os[c] = {ps[c].type_tag, ps[c].nullable};
    }
    const uint32_t block_size = pick(5) ? uint32_t(pick(70000)) : interesting(65516, 65516);
    uint32_t max_rows = pick(4) ? uint32_t(pick(10000)) : interesting(1000, block_size);
    if (ncols && ncols <= 64 && pick(50) == 0) {  // one unknown tag, modest row count: InvalidTypeTag on both sides
      const uint32_t c = uint32_t(pick(ncols));
      ps[c].type_tag = os[c].type_tag = uint16_t(pick(2) ? 0 : 11 + pick(1000));
      max_rows = uint32_t(pick(1000));
    }
    pgf_layout_plan pp;
    orc_layout_plan op;
    std::memset(&pp, 0, sizeof pp);
    std::memset(&op, 0, sizeof op);
    const int a = pgf::plan_layout(ps.data(), ncols, max_rows, block_size, &pp);
    const int b = orc_layout_plan_new(os.data(), ncols, max_rows, block_size, &op);
    bool same = a == b;
    if (same && a == PGF_OK) {
      same = pp.block_size == op.block_size && pp.max_rows == op.max_rows && pp.front_base == op.front_base && pp.pool_base == op.pool_base && pp.ncols == op.ncols;
      for (uint32_t c = 0; same && c < ncols; ++c)
        same = pp.cols[c].type_tag == op.cols[c].type_tag && pp.cols[c].flags == op.cols[c].flags && pp.cols[c].validity_off == op.cols[c].validity_off &&
               pp.cols[c].values_off == op.cols[c].values_off && pp.cols[c].validity_len == op.cols[c].validity_len && pp.cols[c].values_len == op.cols[c].values_len;
    }
    uint32_t cap_p = 0, cap_o = 0;
    if (it % 16 == 0) {  // the cap search plans ~20 times: sample it
      const int ca = pgf::fixed_row_cap(ps.data(), ncols, block_size, &cap_p);
      const int cb = orc_fixed_row_cap(os.data(), ncols, block_size, &cap_o);
      same = same && ca == cb && (ca != PGF_OK || cap_p == cap_o);
    }
    if (!same) {
      if (disagreements < 10) std::printf("PLAN DISAGREE ncols=%u max_rows=%u block_size=%u status %d/%d cap %u/%u\n", ncols, max_rows, block_size, a, b, cap_p, cap_o);
      ++disagreements;
    }
    ++plans;
  }
  // ---- transfer page header: decode of arbitrary 20 bytes, and encode/decode round trips
  uint64_t headers = 0;
  for (uint64_t it = 0; it < iterations; ++it) {
    uint8_t h[20];
    const uint16_t kind = uint16_t(rng()), flags = uint16_t(rng());
    const uint32_t payload = uint32_t(rng());
    pgf::encode_page_header(kind, flags, payload, h);
    uint8_t ho[20];
    orc_page_header_encode(kind, flags, payload, ho);
    bool same = std::memcmp(h, ho, 20) == 0;
    const uint64_t nmut = pick(3);
    for (uint64_t k = 0; k < nmut; ++k) h[pick(20)] = pick(2) ? uint8_t(rng()) : uint8_t(h[pick(20)] ^ (1u << pick(8)));
    uint16_t k1 = 0, f1 = 0, k2 = 0, f2 = 0;
    uint32_t p1 = 0, p2 = 0;
    const int a = pgf::decode_page_header(h, &k1, &f1, &p1);
    const int b = orc_page_header_decode(h, &k2, &f2, &p2);
    same = same && a == b && (a != PGF_OK || (k1 == k2 && f1 == f2 && p1 == p2));
    if (nmut == 0) same = same && a == PGF_OK && k1 == kind && f1 == flags && p1 == payload;
    if (!same) {
      if (disagreements < 10) std::printf("HEADER DISAGREE status %d/%d\n", a, b);
      ++disagreements;
    }
    ++headers;
  }
  std::printf("plans %llu headers %llu rowwise_blocks %llu\n", (unsigned long long)plans, (unsigned long long)headers, (unsigned long long)rowwise_blocks);
  std::printf("codes");
  for (int c = 0; c < 256; ++c)
    if (histogram[c]) std::printf(" %d:%llu", c, (unsigned long long)histogram[c]);
  std::printf("\n");
  std::printf("iterations %llu accepted %llu rejected %llu disagreements %llu\n", (unsigned long long)iterations,
              (unsigned long long)accepted, (unsigned long long)rejected, (unsigned long long)disagreements);
  return disagreements ? 1 : 0;
}
