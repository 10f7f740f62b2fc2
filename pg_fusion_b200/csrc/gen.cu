// Counter-based synthetic TPC-H-shaped data generator (row id -> values), writing pages in
// the reference's exact on-page format (page/transfer header + page/arrow_layout block)
// directly into HBM.  Shapes follow the reference's own TPC-H harness
// (benches/tpch/schema.sql:60-89: money as Float64, dates as ISO text => inline Utf8View,
// keys as Int32) and SURVEY.md section 8d.
#include "bloom_device.cuh"
#include "context.hpp"
#include "layout.hpp"

namespace pgf {

namespace {

constexpr int kGenThreads = 256;

struct GenParams {
  uint8_t* pages;
  uint64_t page_stride;
  uint32_t block_size;
  uint32_t ncols;
  uint32_t rows_per_page;
  uint32_t npages;
  uint64_t first_row, rows, seed, scale_rows;
  int32_t table;
  int32_t dense_keys;
  uint32_t front_base, pool_base;
  uint32_t values_off[kMaxStageCols], validity_off[kMaxStageCols];
  uint16_t type_tag[kMaxStageCols];
  uint8_t page_header[20];
};

// one independent 64-bit random value per (row, stream)
__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t row, uint32_t stream) {
  return splitmix64(splitmix64(seed + 0x9E3779B97F4A7C15ull * (row + 1)) ^ (0xD6E8FEB86659FD93ull * (stream + 1)));
}

// days since 1992-01-01 -> "YYYY-MM-DD" inline view (len 10, zero padded)
__device__ __forceinline__ uint4 date_view(uint32_t days_from_1992) {
  // civil-from-days (proleptic Gregorian); 1992-01-01 is day 8035 since 1970-01-01
  const int64_t z = int64_t(days_from_1992) + 8035 + 719468;
  const int64_t era = z / 146097;
  const uint32_t doe = uint32_t(z - era * 146097);
  const uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const uint32_t y0 = yoe + uint32_t(era) * 400;
  const uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const uint32_t mp = (5 * doy + 2) / 153;
  const uint32_t d = doy - (153 * mp + 2) / 5 + 1;
  const uint32_t m = mp < 10 ? mp + 3 : mp - 9;
  const uint32_t y = y0 + (m <= 2);
  uint8_t s[12] = {uint8_t('0' + y / 1000), uint8_t('0' + y / 100 % 10), uint8_t('0' + y / 10 % 10), uint8_t('0' + y % 10), '-',
                   uint8_t('0' + m / 10), uint8_t('0' + m % 10), '-', uint8_t('0' + d / 10), uint8_t('0' + d % 10), 0, 0};
  uint4 v;
  v.x = 10;
  memcpy(&v.y, s, 12);
  return v;
}

__device__ __forceinline__ uint4 str_view(const char* s, uint32_t n) {
  uint8_t b[12] = {0};
  for (uint32_t i = 0; i < n; ++i) b[i] = uint8_t(s[i]);
  uint4 v;
  v.x = n;
  memcpy(&v.y, b, 12);
  return v;
}

// TPC-H sparse order keys: the first 8 of every 32 key values are used
__device__ __forceinline__ int32_t order_key(uint64_t order_index) { return int32_t((order_index / 8) * 32 + order_index % 8 + 1); }
__device__ __forceinline__ uint32_t order_date_days(uint64_t seed, uint64_t order_index) {
  return uint32_t(rnd(seed ^ 0x0BDE5ull, order_index, 1) % 2406);  // 1992-01-01 .. 1998-08-02
}

constexpr uint32_t kCutoffDays = 1263;  // 1995-06-17 as days since 1992-01-01

struct LineItem {
  double quantity, extendedprice, discount, tax;
  int64_t qty_c, price_c, disc_c, tax_c;  // the same values as Decimal128(15,2): unscaled hundredths
  uint32_t ship_days;
  char returnflag, linestatus;
};

__device__ __forceinline__ LineItem lineitem_row(uint64_t seed, uint64_t row) {
  LineItem li;
  const uint32_t qty = 1 + uint32_t(rnd(seed, row, 0) % 50);
  const uint32_t part_cents = 90000 + uint32_t(rnd(seed, row, 1) % 120001);  // 900.00 .. 2100.00
  li.quantity = double(qty);
  li.extendedprice = double(uint64_t(qty) * part_cents) / 100.0;
  li.disc_c = int64_t(rnd(seed, row, 2) % 11);
  li.tax_c = int64_t(rnd(seed, row, 3) % 9);
  li.discount = double(li.disc_c) / 100.0;
  li.tax = double(li.tax_c) / 100.0;
  li.qty_c = int64_t(qty) * 100;
  li.price_c = int64_t(uint64_t(qty) * part_cents);
  li.ship_days = 1 + uint32_t(rnd(seed, row, 4) % 2526);  // 1992-01-02 .. 1998-12-01
  const uint32_t receipt = li.ship_days + 1 + uint32_t(rnd(seed, row, 5) % 30);
  li.linestatus = li.ship_days > kCutoffDays ? 'O' : 'F';
  li.returnflag = receipt <= kCutoffDays ? ((rnd(seed, row, 6) & 1) ? 'R' : 'A') : 'N';
  return li;
}

__global__ void __launch_bounds__(kGenThreads) gen_pages_kernel(const GenParams p) {
  const uint32_t page = blockIdx.x;
  if (page >= p.npages) return;
  uint8_t* base = p.pages + page * p.page_stride;
  uint8_t* block = base + kPageHeaderLen;
  const uint64_t row0 = uint64_t(page) * p.rows_per_page;
  const uint32_t nrows = uint32_t(min(uint64_t(p.rows_per_page), p.rows - row0));
  // zero the front matter (header, descriptors, validity regions are rewritten below)
  for (uint32_t i = threadIdx.x; i < (kPageHeaderLen + p.front_base + 3) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 20; ++i) base[i] = p.page_header[i];
    BlockHeader h{};
    h.magic = kBlockMagic;
    h.version = kBlockVersion;
    h.block_size = p.block_size;
    h.max_rows = p.rows_per_page;
    h.row_count = nrows;
    h.col_count = uint16_t(p.ncols);
    h.front_base = p.front_base;
    h.pool_base = p.pool_base;
    h.tail_cursor = p.block_size;
    memcpy(block, &h, sizeof h);
    for (uint32_t c = 0; c < p.ncols; ++c) {
      ColumnDesc d{};
      d.type_tag = p.type_tag[c];
      d.flags = (p.type_tag[c] == PGF_T_UTF8VIEW || p.type_tag[c] == PGF_T_BINARYVIEW) ? kFlagView : 0;
      d.validity_off = p.validity_off[c];
      d.values_off = p.values_off[c];
      memcpy(block + sizeof(BlockHeader) + c * sizeof(ColumnDesc), &d, sizeof d);
    }
  }
  // validity bitmaps: writers set the bit of every written row (access.rs:316-319)
  const uint32_t vbytes = ((p.rows_per_page + 7) / 8 + 15) & ~15u;
  for (uint32_t c = 0; c < p.ncols; ++c)
    for (uint32_t b = threadIdx.x; b < vbytes; b += blockDim.x) {
      uint8_t v = 0;
      if (b * 8 + 8 <= nrows) v = 0xFF;
      else if (b * 8 < nrows) v = uint8_t((1u << (nrows - b * 8)) - 1);
      block[p.validity_off[c] + b] = v;
    }
  auto f64col = [&](uint32_t c) { return reinterpret_cast<double*>(block + p.values_off[c]); };
  auto i32col = [&](uint32_t c) { return reinterpret_cast<int32_t*>(block + p.values_off[c]); };
  auto viewcol = [&](uint32_t c) { return reinterpret_cast<uint4*>(block + p.values_off[c]); };
  for (uint32_t r = threadIdx.x; r < p.rows_per_page; r += blockDim.x) {
    const bool live = r < nrows;
    const uint64_t row = p.first_row + row0 + r;
    switch (p.table) {
      case PGF_GEN_LINEITEM_Q6: {
        LineItem li{};
        if (live) li = lineitem_row(p.seed, row);
        f64col(0)[r] = li.quantity; f64col(1)[r] = li.extendedprice; f64col(2)[r] = li.discount;
        viewcol(3)[r] = live ? date_view(li.ship_days) : make_uint4(0, 0, 0, 0);
        break;
      }
      case PGF_GEN_LINEITEM_Q1: {
        LineItem li{};
        if (live) li = lineitem_row(p.seed, row);
        f64col(0)[r] = li.quantity; f64col(1)[r] = li.extendedprice; f64col(2)[r] = li.discount; f64col(3)[r] = li.tax;
        viewcol(4)[r] = live ? str_view(&li.returnflag, 1) : make_uint4(0, 0, 0, 0);
        viewcol(5)[r] = live ? str_view(&li.linestatus, 1) : make_uint4(0, 0, 0, 0);
        viewcol(6)[r] = live ? date_view(li.ship_days) : make_uint4(0, 0, 0, 0);
        break;
      }
      case PGF_GEN_LINEITEM_Q3: {
        // each lineitem picks an order uniformly (mean fan-out = lineitems / orders ~ 4);
        // l_shipdate = o_orderdate + 1..121 days as in TPC-H, which drives Q3's selectivity
        const uint64_t norders = p.scale_rows ? p.scale_rows : 1;
        const uint64_t oi = rnd(p.seed, row, 7) % norders;
        LineItem li{};
        if (live) li = lineitem_row(p.seed, row);
        i32col(0)[r] = live ? order_key(oi) : 0;
        f64col(1)[r] = li.extendedprice; f64col(2)[r] = li.discount;
        viewcol(3)[r] = live ? date_view(order_date_days(p.seed, oi) + 1 + uint32_t(rnd(p.seed, row, 8) % 121)) : make_uint4(0, 0, 0, 0);
        break;
      }
      case PGF_GEN_ORDERS_Q3: {
        // o_custkey uniform over the customers with custkey % 3 != 0 (TPC-H)
        const uint64_t ncust = p.scale_rows ? p.scale_rows : 1;
        uint64_t ck = 1 + rnd(p.seed, row, 9) % ncust;
        if (ck % 3 == 0) ck = ck > 1 ? ck - 1 : ck + 1;
        i32col(0)[r] = live ? order_key(row) : 0;
        i32col(1)[r] = live ? int32_t(ck) : 0;
        viewcol(2)[r] = live ? date_view(order_date_days(p.seed, row)) : make_uint4(0, 0, 0, 0);
        i32col(3)[r] = 0;  // o_shippriority is constant 0 in TPC-H
        break;
      }
      case PGF_GEN_CUSTOMER_Q3: {
        const char* seg[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"};
        const uint32_t len[5] = {10, 8, 9, 9, 9};
        const uint32_t k = uint32_t(rnd(p.seed, row, 10) % 5);
        i32col(0)[r] = live ? int32_t(row + 1) : 0;
        viewcol(1)[r] = live ? str_view(seg[k], len[k]) : make_uint4(0, 0, 0, 0);
        break;
      }
      case PGF_GEN_LINEITEM_Q6_D: {
        // the D variant: money Decimal128(15,2) in 16-byte slots, dates Date32 (days since 1970)
        LineItem li{};
        if (live) li = lineitem_row(p.seed, row);
        auto dec = [&](uint32_t c, int64_t v) { reinterpret_cast<longlong2*>(block + p.values_off[c])[r] = make_longlong2(v, v < 0 ? -1 : 0); };
        dec(0, li.qty_c); dec(1, li.price_c); dec(2, li.disc_c);
        i32col(3)[r] = live ? int32_t(li.ship_days + 8035) : 0;
        break;
      }
      case PGF_GEN_LINEITEM_Q1_D: {
        LineItem li{};
        if (live) li = lineitem_row(p.seed, row);
        auto dec = [&](uint32_t c, int64_t v) { reinterpret_cast<longlong2*>(block + p.values_off[c])[r] = make_longlong2(v, v < 0 ? -1 : 0); };
        dec(0, li.qty_c); dec(1, li.price_c); dec(2, li.disc_c); dec(3, li.tax_c);
        reinterpret_cast<int16_t*>(block + p.values_off[4])[r] = live ? int16_t(li.returnflag) : 0;   // flag codes
        reinterpret_cast<int16_t*>(block + p.values_off[5])[r] = live ? int16_t(li.linestatus) : 0;
        i32col(6)[r] = live ? int32_t(li.ship_days + 8035) : 0;
        break;
      }
      default: {  // PGF_GEN_KEYS_I64
        int64_t* col = reinterpret_cast<int64_t*>(block + p.values_off[0]);
        col[r] = live ? (p.dense_keys ? int64_t(row + 1) : int64_t(splitmix64(p.seed + row))) : 0;
        break;
      }
    }
  }
}

}  // namespace

pgf_status gen_schema(int32_t table, pgf_column_spec* schema, uint32_t* ncols) {
  auto set = [&](std::initializer_list<int> types) {
    uint32_t n = 0;
    for (int t : types) schema[n++] = pgf_column_spec{uint16_t(t), 0};  // NOT NULL columns (schema.sql)
    *ncols = n;
  };
  switch (table) {
    case PGF_GEN_LINEITEM_Q6: set({PGF_T_FLOAT64, PGF_T_FLOAT64, PGF_T_FLOAT64, PGF_T_UTF8VIEW}); break;
    case PGF_GEN_LINEITEM_Q1: set({PGF_T_FLOAT64, PGF_T_FLOAT64, PGF_T_FLOAT64, PGF_T_FLOAT64, PGF_T_UTF8VIEW, PGF_T_UTF8VIEW, PGF_T_UTF8VIEW}); break;
    case PGF_GEN_LINEITEM_Q3: set({PGF_T_INT32, PGF_T_FLOAT64, PGF_T_FLOAT64, PGF_T_UTF8VIEW}); break;
    case PGF_GEN_ORDERS_Q3: set({PGF_T_INT32, PGF_T_INT32, PGF_T_UTF8VIEW, PGF_T_INT32}); break;
    case PGF_GEN_CUSTOMER_Q3: set({PGF_T_INT32, PGF_T_UTF8VIEW}); break;
    case PGF_GEN_KEYS_I64: set({PGF_T_INT64}); break;
    case PGF_GEN_LINEITEM_Q6_D: set({PGF_T_DECIMAL128, PGF_T_DECIMAL128, PGF_T_DECIMAL128, PGF_T_INT32}); break;
    case PGF_GEN_LINEITEM_Q1_D: set({PGF_T_DECIMAL128, PGF_T_DECIMAL128, PGF_T_DECIMAL128, PGF_T_DECIMAL128, PGF_T_INT16, PGF_T_INT16, PGF_T_INT32}); break;
    default: return PGF_ERR_INVALID_ARGUMENT;
  }
  return PGF_OK;
}

pgf_status gen_scan(pgf_ctx* ctx, uint64_t scan_id, const pgf_gen_spec* spec) {
  pgf_column_spec schema[PGF_MAX_COLS];
  uint32_t ncols = 0;
  if (gen_schema(spec->table, schema, &ncols)) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "unknown generator table %d", spec->table);
  const uint32_t block_size = ctx->page_size - kPageHeaderLen;
  uint32_t cap = 0;
  PGF_TRY(fixed_row_cap(schema, ncols, block_size, &cap));
  if (cap == 0) return ctx->fail(PGF_ERR_LAYOUT_DOES_NOT_FIT, "page too small for the schema");
  pgf_layout_plan plan;
  PGF_TRY(plan_layout(schema, ncols, cap, block_size, &plan));
  const uint64_t npages = (spec->rows + cap - 1) / cap;
  if (npages > 0xFFFFFFF0ull) return ctx->fail(PGF_ERR_INVALID_ARGUMENT, "too many pages");
  PGF_TRY(pgf_scan_declare(ctx, scan_id, schema, ncols, npages ? npages : 1));
  Scan* s;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    s = ctx->scans[scan_id].get();
  }
  CU(ctx, cudaSetDevice(ctx->device));
  GenParams p{};
  p.pages = s->d_pages;
  p.page_stride = ctx->page_size;
  p.block_size = block_size;
  p.ncols = ncols;
  p.rows_per_page = cap;
  p.npages = uint32_t(npages);
  p.first_row = spec->first_row;
  p.rows = spec->rows;
  p.seed = spec->seed;
  p.scale_rows = spec->scale_rows;
  p.table = spec->table;
  p.dense_keys = spec->dense_keys;
  p.front_base = plan.front_base;
  p.pool_base = plan.pool_base;
  LayoutClass lc{};
  lc.max_rows = cap;
  lc.pool_base = plan.pool_base + kPageHeaderLen;
  for (uint32_t c = 0; c < ncols; ++c) {
    p.values_off[c] = plan.cols[c].values_off;
    p.validity_off[c] = plan.cols[c].validity_off;
    p.type_tag[c] = schema[c].type_tag;
    lc.values_off[c] = plan.cols[c].values_off + kPageHeaderLen;
    lc.validity_off[c] = plan.cols[c].validity_off + kPageHeaderLen;
  }
  encode_page_header(PGF_ARROW_LAYOUT_BATCH_KIND, 0, block_size, p.page_header);
  if (npages) {
    gen_pages_kernel<<<uint32_t(npages), kGenThreads, 0, ctx->compute_stream>>>(p);
    CU(ctx, cudaGetLastError());
  }
  // host-side descriptor table (row counts are known without reading the pages back)
  s->h_classes.assign(1, lc);
  s->h_descs.resize(npages);
  for (uint64_t pg = 0; pg < npages; ++pg) {
    PageDesc d{};
    d.row_count = uint32_t(std::min<uint64_t>(cap, spec->rows - pg * cap));
    d.layout_class = 0;
    d.null_mask = 0;
    d.row_base = pg * cap;
    s->h_descs[pg] = d;
  }
  s->npages = npages;
  s->rows = spec->rows;
  s->max_page_rows = npages ? uint32_t(std::min<uint64_t>(cap, spec->rows)) : 0;
  s->descs_dirty = true;
  CU(ctx, cudaStreamSynchronize(ctx->compute_stream));
  return pgf_scan_finish(ctx, scan_id);
}

}  // namespace pgf
