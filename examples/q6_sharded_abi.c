/* Multi-GPU through the C ABI only: one process per GPU, the TPC-H Q6 shape over page shards, the partial aggregate
 * states exchanged and merged INSIDE the library (pgf_comm_*, NCCL over NVLink) -- no Python, no torch, no MPI.
 *   gcc -std=c11 -I include examples/q6_sharded_abi.c -L pg_fusion_b200 -lpgf_b200 -Wl,-rpath,$PWD/pg_fusion_b200 -o q6_sharded
 *   for r in 0 1; do ./q6_sharded $r 2 /tmp/pgf.id 20000000 & done; wait
 * Rank 0 creates the communicator id and publishes it through a file (the worker would use its shared-memory
 * control region); every rank generates its shard of the table on its own GPU (device = rank) and gets the same,
 * merged result from pgf_pipeline_run_sharded.  With world = 1 the call degenerates to pgf_pipeline_run. */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "pgf_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    pgf_status st_ = (call);                                                         \
    if (st_ != PGF_OK) {                                                             \
      fprintf(stderr, "rank %d: %s failed: status %d: %s\n", rank, #call, (int)st_, ctx ? pgf_last_error(ctx) : ""); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

static pgf_pred_term term_str(int col, int cmp, const char *s) {
  pgf_pred_term t;
  memset(&t, 0, sizeof t);
  t.col.col = col; t.cmp = cmp;
  t.lit.type_tag = PGF_T_UTF8VIEW; t.lit.slen = (int32_t)strlen(s);
  memcpy(t.lit.str, s, strlen(s));
  return t;
}
static pgf_pred_term term_f64(int col, int cmp, double v) {
  pgf_pred_term t;
  memset(&t, 0, sizeof t);
  t.col.col = col; t.cmp = cmp;
  t.lit.type_tag = PGF_T_FLOAT64; t.lit.f64 = v;
  return t;
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s rank world id_file [rows]\n", argv[0]); return 2; }
  const int rank = atoi(argv[1]), world = atoi(argv[2]);
  const char *id_file = argv[3];
  const uint64_t rows = argc > 4 ? strtoull(argv[4], NULL, 10) : 4000000ull;
  pgf_ctx *ctx = NULL;
  pgf_config cfg = {rank, 65536, 0, 0};
  CHECK(pgf_ctx_create(&cfg, &ctx));
  if (world > 1) {
    uint8_t id[PGF_COMM_ID_BYTES];
    if (rank == 0) {
      CHECK(pgf_comm_unique_id(id));
      char tmp[512];
      snprintf(tmp, sizeof tmp, "%s.tmp", id_file);
      FILE *f = fopen(tmp, "wb");
      if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) return 2;
      fclose(f);
      if (rename(tmp, id_file)) return 2;
    } else {
      FILE *f = NULL;
      for (int tries = 0; tries < 600 && !(f = fopen(id_file, "rb")); ++tries) nanosleep(&(struct timespec){0, 100000000}, NULL);
      if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "rank %d: no communicator id\n", rank); return 2; }
      fclose(f);
    }
    CHECK(pgf_comm_init(ctx, id, rank, world));
  }
  /* this rank's contiguous shard of the table (pages shard by page, like the reference's CTID-range scan producers) */
  const uint64_t lo = rows * (uint64_t)rank / (uint64_t)world, hi = rows * (uint64_t)(rank + 1) / (uint64_t)world;
  pgf_gen_spec gen = {PGF_GEN_LINEITEM_Q6, 0, 42, lo, hi - lo, 0};
  CHECK(pgf_gen_scan(ctx, 1, &gen));
  pgf_pipeline *plan = calloc(1, sizeof *plan);
  plan->scan_id = 1;
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
  pgf_result *res = NULL;
  CHECK(pgf_pipeline_run_sharded(ctx, plan, 1, &res));
  printf("rank %d of %d: shard rows=%llu merged revenue=%.6f count=%lld\n", rank, world, (unsigned long long)(hi - lo),
         res->aggs[0].f64, (long long)res->aggs[1].lo);
  /* every rank holds the same bits: exchange the result once more, as bytes, and compare */
  int ok = 1;
  if (world > 1) {
    double mine[2] = {res->aggs[0].f64, (double)res->aggs[1].lo}, *all = malloc(sizeof mine * (size_t)world);
    CHECK(pgf_comm_all_gather_host(ctx, mine, all, sizeof mine));
    for (int r = 0; r < world; ++r) ok &= memcmp(all + 2 * r, mine, sizeof mine) == 0;
    free(all);
  }
  pgf_result_free(res);
  free(plan);
  CHECK(pgf_scan_release(ctx, 1));
  if (world > 1) CHECK(pgf_comm_destroy(ctx));
  pgf_ctx_destroy(ctx);
  puts(ok ? "ok" : "MISMATCH");
  return ok ? 0 : 3;
}
