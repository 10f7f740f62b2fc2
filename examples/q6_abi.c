/* Plain-C consumer of include/pgf_b200.h: the TPC-H Q6 shape through the C ABI only, the way the
 * Rust shim of INTEGRATION.md drives the library (declare scan -> push pages -> finish -> run).
 *   gcc -std=c11 -I include examples/q6_abi.c -L pg_fusion_b200 -lpgf_b200 -Wl,-rpath,$PWD/pg_fusion_b200 -o q6_abi
 *   ./q6_abi [rows]
 * Pages are produced by the library's device generator, read back to the host and pushed again
 * through pgf_scan_push_pages, so the ingest path (host admission + H2D + device import checks)
 * is exercised exactly like a stream of shared-memory pages would. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pgf_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    pgf_status st_ = (call);                                                         \
    if (st_ != PGF_OK) {                                                             \
      fprintf(stderr, "%s failed: status %d: %s\n", #call, (int)st_, ctx ? pgf_last_error(ctx) : ""); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static pgf_pred_term term_str(int col, int cmp, const char *s) {
  pgf_pred_term t;
  memset(&t, 0, sizeof t);
  t.col.source = 0; t.col.col = col; t.cmp = cmp;
  t.lit.type_tag = PGF_T_UTF8VIEW; t.lit.slen = (int32_t)strlen(s);
  memcpy(t.lit.str, s, strlen(s));
  return t;
}
static pgf_pred_term term_f64(int col, int cmp, double v) {
  pgf_pred_term t;
  memset(&t, 0, sizeof t);
  t.col.source = 0; t.col.col = col; t.cmp = cmp;
  t.lit.type_tag = PGF_T_FLOAT64; t.lit.f64 = v;
  return t;
}

int main(int argc, char **argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], NULL, 10) : 1000000ull;
  pgf_ctx *ctx = NULL;
  pgf_config cfg = {0, 65536, 0, 0};
  CHECK(pgf_ctx_create(&cfg, &ctx));

  /* synthetic lineitem pages (l_quantity, l_extendedprice, l_discount f64; l_shipdate utf8view) */
  pgf_gen_spec gen = {PGF_GEN_LINEITEM_Q6, 0, 42, 0, rows, 0};
  CHECK(pgf_gen_scan(ctx, 1, &gen));
  pgf_scan_info info;
  CHECK(pgf_scan_get_info(ctx, 1, &info));
  uint8_t *pages = malloc(info.bytes);
  if (!pages) return 2;
  CHECK(pgf_scan_read_pages(ctx, 1, 0, info.pages, pages));

  /* the scan a worker would feed: declare, push the pages as they arrive, finish */
  pgf_column_spec schema[PGF_MAX_COLS];
  uint32_t ncols = 0;
  CHECK(pgf_gen_schema(PGF_GEN_LINEITEM_Q6, schema, &ncols));
  CHECK(pgf_scan_declare(ctx, 2, schema, ncols, info.pages));
  for (uint64_t p = 0; p < info.pages; p += 256) {
    const uint64_t n = info.pages - p < 256 ? info.pages - p : 256;
    CHECK(pgf_scan_push_pages(ctx, 2, pages + p * 65536, n, 65536));
  }
  CHECK(pgf_scan_finish(ctx, 2));

  /* SELECT sum(l_extendedprice * l_discount), count(*) WHERE l_shipdate >= '1994-01-01' AND l_shipdate < '1995-01-01'
   *   AND l_discount BETWEEN 0.05 AND 0.07 AND l_quantity < 24   (benches/tpch/queries/q06.sql) */
  pgf_pipeline *plan = calloc(1, sizeof *plan);
  plan->scan_id = 2;
  plan->nterms = 5;
  plan->terms[0] = term_str(3, PGF_CMP_GE, "1994-01-01");
  plan->terms[1] = term_str(3, PGF_CMP_LT, "1995-01-01");
  plan->terms[2] = term_f64(2, PGF_CMP_GE, 0.05);
  plan->terms[3] = term_f64(2, PGF_CMP_LE, 0.07);
  plan->terms[4] = term_f64(0, PGF_CMP_LT, 24.0);
  plan->sink = PGF_SINK_AGGREGATE;
  plan->nexprs = 1;
  plan->exprs[0].nfactors = 2;
  plan->exprs[0].factors[0].kind = PGF_FACTOR_COL; plan->exprs[0].factors[0].col.col = 1;
  plan->exprs[0].factors[1].kind = PGF_FACTOR_COL; plan->exprs[0].factors[1].col.col = 2;
  plan->naggs = 2;
  plan->aggs[0].func = PGF_AGG_SUM; plan->aggs[0].expr = 0;
  plan->aggs[1].func = PGF_AGG_COUNT_STAR; plan->aggs[1].expr = -1;
  CHECK(pgf_pipeline_check(ctx, plan));
  pgf_result *res = NULL;
  CHECK(pgf_pipeline_run(ctx, plan, &res));
  printf("rows_in=%llu rows_kept=%llu revenue=%.4f count=%lld kernel_ms=%.4f\n", (unsigned long long)res->rows_in,
         (unsigned long long)res->rows_filtered, res->aggs[0].f64, (long long)res->aggs[1].lo, res->kernel_ms);
  /* the result as a reference result page (what ResultPageProducer would hand to the transport) */
  uint8_t page[65536];
  uint64_t npages = 0, done = 0;
  CHECK(pgf_result_encode_pages(res, 65536, 0, page, 1, &npages, &done));
  printf("result pages=%llu rows=%llu\n", (unsigned long long)npages, (unsigned long long)done);
  const int ok = res->rows_in == rows && (uint64_t)res->aggs[1].lo == res->rows_filtered && npages == 1 && done == 1;
  pgf_result_free(res);
  free(plan);
  free(pages);
  CHECK(pgf_scan_release(ctx, 1));
  CHECK(pgf_scan_release(ctx, 2));
  pgf_ctx_destroy(ctx);
  puts(ok ? "ok" : "MISMATCH");
  return ok ? 0 : 3;
}
