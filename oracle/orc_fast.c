/*
 * oracle/orc_fast.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see orc.h): tight,
 * page-sharded multi-thread loops for the TPC-H Q6 / Q1 shapes over the reference's
 * page format.  These are the "reference-semantics CPU" baselines C1 (1 thread, the
 * reference's actual single-partition execution model, worker_runtime/src/runtime.rs:
 * 748-758) and C2 (all cores, Partial -> Final merge) of BASELINE.md section 3.
 * They are checked against the generic interpreter (orc_ops.c) in tests.
 *
 * Queries: benches/tpch/queries/q06.sql, q01.sql.  Schema: benches/tpch/schema.sql:73-89
 * (money Float64, dates as 10-byte ISO text => inline Utf8View).  [DF-K] semantics as in
 * orc_ops.c; Float64 expressions are evaluated per row exactly as written, no FMA
 * contraction (compile with -ffp-contract=off).
 */
#include "orc.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define HDR 20u

typedef struct {
  const uint8_t *pages;
  uint64_t p0, p1, stride;
  const int32_t *cols;
  /* q6 */
  uint8_t date_lo[12], date_hi[12];
  double disc_lo, disc_hi, qty_lt;
  orc_q6_result q6;
  /* q1 */
  uint8_t date_le[12];
  int with_tax;            /* bit 0: sum_charge with tax; bit 1: compensated (Neumaier) sums */
  orc_q1_result q1;
  double comp[16][5];      /* compensation terms of the per-group sums (bit 1) */
  int rc;
} job;

static inline uint64_t be64(const uint8_t *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return __builtin_bswap64(v);
}
static inline uint32_t be32(const uint8_t *p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return __builtin_bswap32(v);
}

/* bytewise lexicographic compare of an inline view (<= 12 bytes, zero padded,
 * page/arrow_layout/src/raw.rs:114-126) against a zero-padded 12-byte literal of
 * length litlen: compare the padded images as big-endian integers, then the lengths. */
static inline int view_cmp(const orc_byte_view *v, const uint8_t lit[12], int litlen) {
  const uint64_t a = be64(v->data), b = be64(lit);
  if (a != b) return a < b ? -1 : 1;
  const uint32_t c = be32(v->data + 8), d = be32(lit + 8);
  if (c != d) return c < d ? -1 : 1;
  return (v->len > litlen) - (v->len < litlen);
}

static void pad12(uint8_t out[12], const char *s) {
  memset(out, 0, 12);
  memcpy(out, s, strlen(s) > 12 ? 12 : strlen(s));
}

static int page_cols(const uint8_t *page, uint64_t stride, const int32_t *cols, int n,
                     const uint8_t **ptrs, uint32_t *rows) {
  const uint8_t *block = page + HDR;
  orc_block_header h;
  memcpy(&h, block, sizeof h);
  if (h.magic != ORC_BLOCK_MAGIC || (uint64_t)h.block_size + HDR > stride) return -1;
  for (int i = 0; i < n; ++i) {
    if (cols[i] < 0 || cols[i] >= h.col_count) return -1;
    orc_column_desc d;
    memcpy(&d, block + 40 + 20 * (uint32_t)cols[i], sizeof d);
    if (d.null_count != 0) return -1; /* the fast loops handle NOT NULL columns only */
    ptrs[i] = block + d.values_off;
  }
  *rows = h.row_count;
  return 0;
}

static void *q6_worker(void *arg) {
  job *j = arg;
  double sum = 0.0;
  uint64_t kept = 0, rows_in = 0;
  for (uint64_t p = j->p0; p < j->p1; ++p) {
    const uint8_t *ptr[4];
    uint32_t rows;
    if (page_cols(j->pages + p * j->stride, j->stride, j->cols, 4, ptr, &rows)) { j->rc = -1; return NULL; }
    const double *qty = (const double *)ptr[0], *price = (const double *)ptr[1],
                 *disc = (const double *)ptr[2];
    const orc_byte_view *date = (const orc_byte_view *)ptr[3];
    rows_in += rows;
    for (uint32_t r = 0; r < rows; ++r) {
      const int keep = view_cmp(&date[r], j->date_lo, 10) >= 0 &&
                       view_cmp(&date[r], j->date_hi, 10) < 0 && disc[r] >= j->disc_lo &&
                       disc[r] <= j->disc_hi && qty[r] < j->qty_lt;
      if (keep) {
        sum += price[r] * disc[r];
        ++kept;
      }
    }
  }
  j->q6.sum = sum;
  j->q6.rows_in = rows_in;
  j->q6.rows_kept = kept;
  return NULL;
}

static int run_jobs(job *jobs, int nthreads, void *(*fn)(void *)) {
  if (nthreads == 1) {
    fn(&jobs[0]);
    return jobs[0].rc;
  }
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  if (!th) return -1;
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    rc |= jobs[t].rc;
  }
  free(th);
  return rc;
}

static job *make_jobs(const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads,
                      const int32_t *cols) {
  job *jobs = calloc((size_t)nthreads, sizeof(job));
  if (!jobs) return NULL;
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].pages = pages;
    jobs[t].stride = stride;
    jobs[t].cols = cols;
    jobs[t].p0 = npages * (uint64_t)t / (uint64_t)nthreads; /* contiguous page ranges */
    jobs[t].p1 = npages * (uint64_t)(t + 1) / (uint64_t)nthreads;
  }
  return jobs;
}

int orc_q6_pages(const uint8_t *pages, uint64_t npages, uint64_t page_stride, int nthreads,
                 const int32_t cols[4], const char *date_lo, const char *date_hi,
                 double disc_lo, double disc_hi, double qty_lt, orc_q6_result *out) {
  if (nthreads < 1) nthreads = 1;
  job *jobs = make_jobs(pages, npages, page_stride, nthreads, cols);
  if (!jobs) return -1;
  for (int t = 0; t < nthreads; ++t) {
    pad12(jobs[t].date_lo, date_lo);
    pad12(jobs[t].date_hi, date_hi);
    jobs[t].disc_lo = disc_lo;
    jobs[t].disc_hi = disc_hi;
    jobs[t].qty_lt = qty_lt;
  }
  const int rc = run_jobs(jobs, nthreads, q6_worker);
  memset(out, 0, sizeof *out);
  for (int t = 0; t < nthreads; ++t) { /* Final: merge partial states in thread order */
    out->sum += jobs[t].q6.sum;
    out->rows_in += jobs[t].q6.rows_in;
    out->rows_kept += jobs[t].q6.rows_kept;
  }
  free(jobs);
  return rc;
}

static orc_q1_group *q1_group(orc_q1_result *r, char rf, char ls) {
  for (uint32_t g = 0; g < r->ngroups; ++g)
    if (r->groups[g].returnflag == rf && r->groups[g].linestatus == ls) return &r->groups[g];
  if (r->ngroups == 16) return NULL;
  orc_q1_group *g = &r->groups[r->ngroups++];
  memset(g, 0, sizeof *g);
  g->returnflag = rf;
  g->linestatus = ls;
  return g;
}

/* Neumaier's compensated addition: s + c is the running sum with O(1) ulp error whatever the
 * number of terms.  Used by the full-size parity tests, where the plain reference-order sum's
 * own rounding error (~sqrt(n)..n ulps) exceeds the 1e-12 bar that is being checked. */
static inline void nadd(double *s, double *c, double x) {
  const double t = *s + x;
  if (fabs(*s) >= fabs(x)) *c += (*s - t) + x;
  else *c += (x - t) + *s;
  *s = t;
}

static void *q1_worker(void *arg) {
  job *j = arg;
  orc_q1_result *res = &j->q1;
  const int compensated = (j->with_tax & 2) != 0;
  double (*comp)[5] = j->comp;
  for (uint64_t p = j->p0; p < j->p1; ++p) {
    const uint8_t *ptr[7];
    uint32_t rows;
    if (page_cols(j->pages + p * j->stride, j->stride, j->cols, 7, ptr, &rows)) { j->rc = -1; return NULL; }
    const double *qty = (const double *)ptr[0], *price = (const double *)ptr[1],
                 *disc = (const double *)ptr[2], *tax = (const double *)ptr[3];
    const orc_byte_view *rf = (const orc_byte_view *)ptr[4], *ls = (const orc_byte_view *)ptr[5],
                        *date = (const orc_byte_view *)ptr[6];
    res->rows_in += rows;
    for (uint32_t r = 0; r < rows; ++r) {
      if (view_cmp(&date[r], j->date_le, 10) > 0) continue;
      if (rf[r].len != 1 || ls[r].len != 1) { j->rc = -1; return NULL; }
      orc_q1_group *g = q1_group(res, (char)rf[r].data[0], (char)ls[r].data[0]);
      if (!g) { j->rc = -1; return NULL; }
      const double disc_price = price[r] * (1.0 - disc[r]);
      if (compensated) {
        double *c = comp[g - res->groups];
        nadd(&g->sum_qty, &c[0], qty[r]);
        nadd(&g->sum_base_price, &c[1], price[r]);
        nadd(&g->sum_disc_price, &c[2], disc_price);
        if (j->with_tax & 1) nadd(&g->sum_charge, &c[3], disc_price * (1.0 + tax[r]));
        nadd(&g->sum_disc, &c[4], disc[r]);
        g->count += 1;
        continue;
      }
      g->sum_qty += qty[r];
      g->sum_base_price += price[r];
      g->sum_disc_price += disc_price;
      if (j->with_tax & 1) g->sum_charge += disc_price * (1.0 + tax[r]);
      g->sum_disc += disc[r];
      g->count += 1;
    }
  }
  return NULL;
}

int orc_q1_pages(const uint8_t *pages, uint64_t npages, uint64_t page_stride, int nthreads,
                 const int32_t cols[7], const char *date_le, int with_tax, orc_q1_result *out) {
  if (nthreads < 1) nthreads = 1;
  job *jobs = make_jobs(pages, npages, page_stride, nthreads, cols);
  if (!jobs) return -1;
  for (int t = 0; t < nthreads; ++t) {
    pad12(jobs[t].date_le, date_le);
    jobs[t].with_tax = with_tax;
  }
  const int rc = run_jobs(jobs, nthreads, q1_worker);
  memset(out, 0, sizeof *out);
  double comp[16][5];
  memset(comp, 0, sizeof comp);
  for (int t = 0; t < nthreads && rc == 0; ++t) {
    out->rows_in += jobs[t].q1.rows_in;
    for (uint32_t g = 0; g < jobs[t].q1.ngroups; ++g) {
      const orc_q1_group *s = &jobs[t].q1.groups[g];
      orc_q1_group *d = q1_group(out, s->returnflag, s->linestatus);
      if (!d) { free(jobs); return -1; }
      if (with_tax & 2) {  /* compensated merge: partial sums and their compensation terms */
        double *c = comp[d - out->groups];
        const double *sc = jobs[t].comp[g];
        nadd(&d->sum_qty, &c[0], s->sum_qty); nadd(&d->sum_qty, &c[0], sc[0]);
        nadd(&d->sum_base_price, &c[1], s->sum_base_price); nadd(&d->sum_base_price, &c[1], sc[1]);
        nadd(&d->sum_disc_price, &c[2], s->sum_disc_price); nadd(&d->sum_disc_price, &c[2], sc[2]);
        nadd(&d->sum_charge, &c[3], s->sum_charge); nadd(&d->sum_charge, &c[3], sc[3]);
        nadd(&d->sum_disc, &c[4], s->sum_disc); nadd(&d->sum_disc, &c[4], sc[4]);
        d->count += s->count;
        continue;
      }
      d->sum_qty += s->sum_qty;
      d->sum_base_price += s->sum_base_price;
      d->sum_disc_price += s->sum_disc_price;
      d->sum_charge += s->sum_charge;
      d->sum_disc += s->sum_disc;
      d->count += s->count;
    }
  }
  if (with_tax & 2)
    for (uint32_t g = 0; g < out->ngroups; ++g) {
      out->groups[g].sum_qty += comp[g][0];
      out->groups[g].sum_base_price += comp[g][1];
      out->groups[g].sum_disc_price += comp[g][2];
      out->groups[g].sum_charge += comp[g][3];
      out->groups[g].sum_disc += comp[g][4];
    }
  free(jobs);
  return rc;
}
