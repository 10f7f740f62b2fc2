/*
 * oracle/orc_q3.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see orc.h): the TPC-H Q3 shape
 * (benches/tpch/queries/q03.sql over benches/tpch/schema.sql:73-89) as tight, page-sharded
 * multi-thread loops over the reference's page format, fed shard by shard so that SF100 can be
 * walked without ever holding a whole table on the host:
 *
 *   customer (c_custkey i32, c_mktsegment text)        WHERE c_mktsegment = :segment
 *     |><| orders (o_orderkey, o_custkey i32, o_orderdate text, o_shippriority i32)
 *                                                       WHERE o_orderdate < :date
 *     |><| lineitem (l_orderkey i32, l_extendedprice, l_discount f64, l_shipdate text)
 *                                                       WHERE l_shipdate > :date
 *   GROUP BY l_orderkey, o_orderdate, o_shippriority ; SUM(l_extendedprice * (1 - l_discount))
 *
 * [DF-K] semantics as in orc_ops.c (HashJoinExec CollectLeft / Inner: duplicates on either side
 * multiply, the plan is the one of SURVEY.md 3.5); checked against the generic interpreter in
 * tests/test_oracle_ops.py.  With nthreads == 1 the per-group Float64 sums are accumulated in input
 * row order, like AggregateExec(mode=Single); with more threads the order inside a group depends on
 * the page split (bench baseline only).  Float64 per row exactly as written (-ffp-contract=off).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "orc.h"

#define HDR 20u

typedef struct {
  int32_t key, prio;
  uint8_t date[12];
  int32_t datelen;
} q3_order;

struct orc_q3 {
  uint8_t segment[12], date[12];
  int seglen, datelen;
  /* customer side */
  int32_t *cust_keys;
  uint64_t ncust, cust_cap;
  int32_t *cust_set;  /* open addressing over distinct keys */
  uint32_t *cust_used; /* 0 = empty slot, else how many build rows carry the key (duplicates multiply) */
  uint64_t cust_mask;
  /* orders side */
  q3_order *orders;
  uint64_t nord, ord_cap;
  uint32_t *ord_slot; /* table of indices into orders[] + 1; 0 = empty */
  uint64_t ord_mask;
  /* per orders row: aggregate state */
  double *sum;
  uint64_t *cnt;
  /* stats */
  uint64_t li_rows, li_filtered, li_joined;
  pthread_mutex_t mu;
};

static inline uint64_t be64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return __builtin_bswap64(v); }
static inline uint32_t be32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return __builtin_bswap32(v); }
static inline int view_cmp(const orc_byte_view *v, const uint8_t lit[12], int litlen) {
  const uint64_t a = be64(v->data), b = be64(lit);
  if (a != b) return a < b ? -1 : 1;
  const uint32_t c = be32(v->data + 8), d = be32(lit + 8);
  if (c != d) return c < d ? -1 : 1;
  return (v->len > litlen) - (v->len < litlen);
}
static inline uint64_t hash32(int32_t k) {
  uint64_t x = (uint32_t)k;
  x *= 0x9E3779B97F4A7C15ull;
  return x ^ (x >> 29);
}

static int page_cols(const uint8_t *page, uint64_t stride, int n, const uint8_t **ptrs, uint32_t *rows) {
  const uint8_t *block = page + HDR;
  orc_block_header h;
  memcpy(&h, block, sizeof h);
  if (h.magic != ORC_BLOCK_MAGIC || (uint64_t)h.block_size + HDR > stride || h.col_count < n) return -1;
  for (int i = 0; i < n; ++i) {
    orc_column_desc d;
    memcpy(&d, block + 40 + 20 * (uint32_t)i, sizeof d);
    if (d.null_count != 0) return -1; /* the fast loops handle NOT NULL columns only */
    ptrs[i] = block + d.values_off;
  }
  *rows = h.row_count;
  return 0;
}

orc_q3 *orc_q3_new(const char *segment, const char *date) {
  orc_q3 *q = calloc(1, sizeof *q);
  if (!q) return NULL;
  q->seglen = (int)strlen(segment);
  q->datelen = (int)strlen(date);
  if (q->seglen > 12 || q->datelen > 12) { free(q); return NULL; }
  memcpy(q->segment, segment, (size_t)q->seglen);
  memcpy(q->date, date, (size_t)q->datelen);
  pthread_mutex_init(&q->mu, NULL);
  return q;
}

void orc_q3_free(orc_q3 *q) {
  if (!q) return;
  free(q->cust_keys); free(q->cust_set); free(q->cust_used);
  free(q->orders); free(q->ord_slot); free(q->sum); free(q->cnt);
  pthread_mutex_destroy(&q->mu);
  free(q);
}

typedef struct {
  orc_q3 *q;
  const uint8_t *pages;
  uint64_t p0, p1, stride;
  int what; /* 0 customer, 1 orders, 2 lineitem, 3 orders table build */
  uint64_t r0, r1;
  /* thread-local outputs */
  int32_t *keys; uint64_t nkeys, keys_cap;
  q3_order *ords; uint64_t nords, ords_cap;
  uint64_t rows, filtered, joined;
  int rc;
} q3_job;

static uint32_t cust_matches(const orc_q3 *q, int32_t k) {
  if (!q->cust_set) return 0;
  for (uint64_t i = hash32(k) & q->cust_mask;; i = (i + 1) & q->cust_mask) {
    if (!q->cust_used[i]) return 0;
    if (q->cust_set[i] == k) return q->cust_used[i];
  }
}

static void *q3_worker(void *arg) {
  q3_job *j = arg;
  orc_q3 *q = j->q;
  if (j->what == 3) { /* parallel build of the orders table: CAS on the slot word */
    for (uint64_t r = j->r0; r < j->r1; ++r) {
      for (uint64_t i = hash32(q->orders[r].key) & q->ord_mask;; i = (i + 1) & q->ord_mask) {
        uint32_t expect = 0;
        if (__atomic_compare_exchange_n(&q->ord_slot[i], &expect, (uint32_t)(r + 1), 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
      }
    }
    return NULL;
  }
  for (uint64_t p = j->p0; p < j->p1; ++p) {
    const uint8_t *ptr[4];
    uint32_t rows;
    const int ncols = j->what == 0 ? 2 : 4;
    if (page_cols(j->pages + p * j->stride, j->stride, ncols, ptr, &rows)) { j->rc = -1; return NULL; }
    j->rows += rows;
    if (j->what == 0) {
      const int32_t *ck = (const int32_t *)ptr[0];
      const orc_byte_view *seg = (const orc_byte_view *)ptr[1];
      for (uint32_t r = 0; r < rows; ++r) {
        if (view_cmp(&seg[r], q->segment, q->seglen) != 0) continue;
        if (j->nkeys == j->keys_cap) {
          j->keys_cap = j->keys_cap ? j->keys_cap * 2 : 4096;
          j->keys = realloc(j->keys, j->keys_cap * sizeof(int32_t));
          if (!j->keys) { j->rc = -1; return NULL; }
        }
        j->keys[j->nkeys++] = ck[r];
      }
    } else if (j->what == 1) {
      const int32_t *ok = (const int32_t *)ptr[0], *ck = (const int32_t *)ptr[1], *prio = (const int32_t *)ptr[3];
      const orc_byte_view *od = (const orc_byte_view *)ptr[2];
      for (uint32_t r = 0; r < rows; ++r) {
        if (view_cmp(&od[r], q->date, q->datelen) >= 0) continue;
        ++j->filtered;
        if (od[r].len > 12) { j->rc = -1; return NULL; }
        for (uint32_t m = cust_matches(q, ck[r]); m; --m) { /* one output row per matching customer row */
          if (j->nords == j->ords_cap) {
            j->ords_cap = j->ords_cap ? j->ords_cap * 2 : 4096;
            j->ords = realloc(j->ords, j->ords_cap * sizeof(q3_order));
            if (!j->ords) { j->rc = -1; return NULL; }
          }
          q3_order *o = &j->ords[j->nords++];
          o->key = ok[r];
          o->prio = prio[r];
          memcpy(o->date, od[r].data, 12);
          o->datelen = od[r].len;
        }
      }
    } else {
      const int32_t *ok = (const int32_t *)ptr[0];
      const double *price = (const double *)ptr[1], *disc = (const double *)ptr[2];
      const orc_byte_view *sd = (const orc_byte_view *)ptr[3];
      for (uint32_t r = 0; r < rows; ++r) {
        if (view_cmp(&sd[r], q->date, q->datelen) <= 0) continue;
        ++j->filtered;
        if (!q->ord_slot) continue;
        const int32_t k = ok[r];
        for (uint64_t i = hash32(k) & q->ord_mask;; i = (i + 1) & q->ord_mask) {
          const uint32_t s = q->ord_slot[i];
          if (!s) break;
          if (q->orders[s - 1].key != k) continue;
          /* one output row per matching build row (duplicates multiply) */
          const double v = price[r] * (1.0 - disc[r]);
          double *dst = &q->sum[s - 1];
          uint64_t old, neu;
          memcpy(&old, dst, 8);
          for (;;) { /* atomic Float64 add (single-threaded runs take the first try) */
            double d;
            memcpy(&d, &old, 8);
            d += v;
            memcpy(&neu, &d, 8);
            if (__atomic_compare_exchange_n((uint64_t *)dst, &old, neu, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
          }
          __atomic_fetch_add(&q->cnt[s - 1], 1, __ATOMIC_RELAXED);
          ++j->joined;
        }
      }
    }
  }
  return NULL;
}

static int run_pages(orc_q3 *q, int what, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads, q3_job **jobs_out) {
  if (nthreads < 1) nthreads = 1;
  q3_job *jobs = calloc((size_t)nthreads, sizeof *jobs);
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  if (!jobs || !th) { free(jobs); free(th); return -1; }
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].q = q; jobs[t].pages = pages; jobs[t].stride = stride; jobs[t].what = what;
    jobs[t].p0 = npages * (uint64_t)t / (uint64_t)nthreads;
    jobs[t].p1 = npages * (uint64_t)(t + 1) / (uint64_t)nthreads;
  }
  if (nthreads == 1) q3_worker(&jobs[0]);
  else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, q3_worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  free(th);
  int rc = 0;
  for (int t = 0; t < nthreads; ++t) rc |= jobs[t].rc;
  *jobs_out = jobs;
  (void)nthreads;
  return rc;
}

int orc_q3_customer(orc_q3 *q, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  q3_job *jobs = NULL;
  int rc = run_pages(q, 0, pages, npages, stride, nthreads, &jobs);
  for (int t = 0; t < nthreads && jobs; ++t) { /* thread order == page order */
    if (!rc && jobs[t].nkeys) {
      if (q->ncust + jobs[t].nkeys > q->cust_cap) {
        q->cust_cap = (q->ncust + jobs[t].nkeys) * 2;
        q->cust_keys = realloc(q->cust_keys, q->cust_cap * sizeof(int32_t));
        if (!q->cust_keys) rc = -1;
      }
      if (!rc) { memcpy(q->cust_keys + q->ncust, jobs[t].keys, jobs[t].nkeys * sizeof(int32_t)); q->ncust += jobs[t].nkeys; }
    }
    free(jobs[t].keys);
  }
  free(jobs);
  return rc;
}

int orc_q3_customer_finish(orc_q3 *q) {
  uint64_t cap = 1024;
  while (cap < q->ncust * 2) cap <<= 1;
  q->cust_mask = cap - 1;
  q->cust_set = malloc(cap * sizeof(int32_t));
  q->cust_used = calloc(cap, sizeof(uint32_t));
  if (!q->cust_set || !q->cust_used) return -1;
  for (uint64_t r = 0; r < q->ncust; ++r) {
    uint64_t i = hash32(q->cust_keys[r]) & q->cust_mask;
    while (q->cust_used[i] && q->cust_set[i] != q->cust_keys[r]) i = (i + 1) & q->cust_mask;
    q->cust_used[i] += 1;
    q->cust_set[i] = q->cust_keys[r];
  }
  return 0;
}

int orc_q3_orders(orc_q3 *q, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  q3_job *jobs = NULL;
  int rc = run_pages(q, 1, pages, npages, stride, nthreads, &jobs);
  for (int t = 0; t < nthreads && jobs; ++t) {
    if (!rc && jobs[t].nords) {
      if (q->nord + jobs[t].nords > q->ord_cap) {
        q->ord_cap = (q->nord + jobs[t].nords) * 2;
        q->orders = realloc(q->orders, q->ord_cap * sizeof(q3_order));
        if (!q->orders) rc = -1;
      }
      if (!rc) { memcpy(q->orders + q->nord, jobs[t].ords, jobs[t].nords * sizeof(q3_order)); q->nord += jobs[t].nords; }
    }
    free(jobs[t].ords);
  }
  free(jobs);
  return rc;
}

int orc_q3_orders_finish(orc_q3 *q, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (q->nord >= 0xFFFFFFFEull) return -1;
  uint64_t cap = 1024;
  while (cap < q->nord * 2) cap <<= 1;
  q->ord_mask = cap - 1;
  q->ord_slot = calloc(cap, sizeof(uint32_t));
  q->sum = calloc(q->nord ? q->nord : 1, sizeof(double));
  q->cnt = calloc(q->nord ? q->nord : 1, sizeof(uint64_t));
  if (!q->ord_slot || !q->sum || !q->cnt) return -1;
  q3_job *jobs = calloc((size_t)nthreads, sizeof *jobs);
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
  if (!jobs || !th) { free(jobs); free(th); return -1; }
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].q = q; jobs[t].what = 3;
    jobs[t].r0 = q->nord * (uint64_t)t / (uint64_t)nthreads;
    jobs[t].r1 = q->nord * (uint64_t)(t + 1) / (uint64_t)nthreads;
  }
  if (nthreads == 1) q3_worker(&jobs[0]);
  else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, q3_worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  free(th);
  free(jobs);
  return 0;
}

int orc_q3_lineitem(orc_q3 *q, const uint8_t *pages, uint64_t npages, uint64_t stride, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  q3_job *jobs = NULL;
  const int rc = run_pages(q, 2, pages, npages, stride, nthreads, &jobs);
  for (int t = 0; t < nthreads && jobs; ++t) {
    q->li_rows += jobs[t].rows;
    q->li_filtered += jobs[t].filtered;
    q->li_joined += jobs[t].joined;
  }
  free(jobs);
  return rc;
}

/* out: customers kept, orders built, lineitem rows in, after the date filter, joined rows, build rows with >= 1 match */
int orc_q3_stats(const orc_q3 *q, uint64_t out[6]) {
  uint64_t groups = 0;
  for (uint64_t r = 0; r < q->nord; ++r) groups += q->cnt && q->cnt[r] != 0;
  out[0] = q->ncust; out[1] = q->nord; out[2] = q->li_rows; out[3] = q->li_filtered; out[4] = q->li_joined; out[5] = groups;
  return 0;
}

/* One record per build row that found a partner (build rows with equal (key, date, priority) are one GROUP BY group:
 * the caller adds them up; TPC-H order keys are unique).  Returns the number of records written. */
uint64_t orc_q3_groups(const orc_q3 *q, int32_t *keys, uint8_t *dates12, int32_t *datelens, int32_t *prios, double *sums,
                       uint64_t *counts, uint64_t cap) {
  uint64_t n = 0;
  for (uint64_t r = 0; r < q->nord && q->cnt; ++r) {
    if (!q->cnt[r]) continue;
    if (n < cap) {
      keys[n] = q->orders[r].key;
      memcpy(dates12 + 12 * n, q->orders[r].date, 12);
      datelens[n] = q->orders[r].datelen;
      prios[n] = q->orders[r].prio;
      sums[n] = q->sum[r];
      counts[n] = q->cnt[r];
    }
    ++n;
  }
  return n;
}
