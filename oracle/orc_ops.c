/*
 * oracle/orc_ops.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see orc.h): scalar,
 * single-thread restatement of the DataFusion 44 / arrow-rs 53.4.1 operators that the
 * reference's worker plans over page-backed batches (worker_runtime/src/runtime.rs:
 * 667-698; executed at pg/extension/src/worker.rs:292):
 *   FilterExec, ProjectionExec/BinaryExpr, AggregateExec (no-group and grouped),
 *   HashJoinExec(CollectLeft, Inner).
 *
 * [DF-K] -- the algorithm lives in un-vendored third-party crates (datafusion = "44.0",
 * arrow-* = "53.4.1", Cargo.toml:49-60).  PARITY UNPINNED at the operator boundary: the
 * reference holds no operator-level golden vectors; this file
 * restates the published semantics listed in SURVEY.md section 8c and is cross-checked
 * against pyarrow/Acero in tests/test_oracle_ops.py and against the reference's
 * PostgreSQL-bound smoke values (pg/extension/src/smoke_tests.rs:205-304).
 */
#include "orc.h"

#include <stdlib.h>
#include <string.h>

typedef __int128 i128;
typedef unsigned __int128 u128;

/* evaluator value */
typedef struct {
  int kind;  /* ORC_V_* */
  int width; /* integer width in bits (16/32/64) for ORC_V_I64 */
  double f;
  i128 i;
  const uint8_t *s;
  uint32_t slen;
  int b;
} val;

typedef struct {
  const orc_table *src[8];
  uint64_t row[8];
} rowctx;

static int load_col(const rowctx *ctx, int source, int colidx, val *out) {
  const orc_table *t = ctx->src[source];
  if (!t || (uint32_t)colidx >= t->ncols) return -1;
  const orc_column *c = &t->cols[colidx];
  const uint64_t r = ctx->row[source];
  memset(out, 0, sizeof *out);
  if (c->validity && !c->validity[r]) { out->kind = ORC_V_NULL; return 0; }
  switch (c->type_tag) {
    case ORC_T_BOOLEAN: out->kind = ORC_V_BOOL; out->b = ((const uint8_t *)c->values)[r]; break;
    case ORC_T_INT16: out->kind = ORC_V_I64; out->width = 16; out->i = ((const int16_t *)c->values)[r]; break;
    case ORC_T_INT32: out->kind = ORC_V_I64; out->width = 32; out->i = ((const int32_t *)c->values)[r]; break;
    case ORC_T_INT64: out->kind = ORC_V_I64; out->width = 64; out->i = ((const int64_t *)c->values)[r]; break;
    case ORC_T_FLOAT32: out->kind = ORC_V_F64; out->f = ((const float *)c->values)[r]; break;
    case ORC_T_FLOAT64: out->kind = ORC_V_F64; out->f = ((const double *)c->values)[r]; break;
    case ORC_T_DECIMAL128: {
      i128 v;
      memcpy(&v, (const uint8_t *)c->values + r * 16, 16);
      out->kind = ORC_V_I128; out->i = v; break;
    }
    case ORC_T_UTF8VIEW: case ORC_T_BINARYVIEW: {
      const orc_str *sv = (const orc_str *)c->values + r;
      out->kind = ORC_V_STR; out->s = sv->ptr; out->slen = sv->len; break;
    }
    default: return -1;
  }
  return 0;
}

static i128 wrap_int(i128 v, int width) {
  switch (width) { /* arrow add_wrapping/sub_wrapping/mul_wrapping at the operand width */
    case 16: return (i128)(int16_t)(uint16_t)(u128)v;
    case 32: return (i128)(int32_t)(uint32_t)(u128)v;
    default: return (i128)(int64_t)(uint64_t)(u128)v;
  }
}

/* f64 total order (arrow-rs comparison kernels for floats) [DF-K] */
static int64_t f64_total_key(double d) {
  int64_t b;
  memcpy(&b, &d, 8);
  b ^= (int64_t)(((uint64_t)(b >> 63)) >> 1);
  return b;
}

static int cmp_vals(const val *a, const val *b, int *out) {
  if (a->kind == ORC_V_F64 || b->kind == ORC_V_F64) {
    /* numeric coercion to Float64 when either side is Float64 */
    double x = a->kind == ORC_V_F64 ? a->f : (double)(int64_t)a->i;
    double y = b->kind == ORC_V_F64 ? b->f : (double)(int64_t)b->i;
    if ((a->kind != ORC_V_F64 && a->kind != ORC_V_I64) || (b->kind != ORC_V_F64 && b->kind != ORC_V_I64)) return -1;
    const int64_t kx = f64_total_key(x), ky = f64_total_key(y);
    *out = (kx > ky) - (kx < ky);
    return 0;
  }
  if ((a->kind == ORC_V_I64 || a->kind == ORC_V_I128) && (b->kind == ORC_V_I64 || b->kind == ORC_V_I128)) {
    *out = (a->i > b->i) - (a->i < b->i);
    return 0;
  }
  if (a->kind == ORC_V_STR && b->kind == ORC_V_STR) {
    /* bytewise lexicographic, shorter-is-less on a common prefix */
    const uint32_t n = a->slen < b->slen ? a->slen : b->slen;
    int c = n ? memcmp(a->s, b->s, n) : 0;
    if (c == 0) c = (a->slen > b->slen) - (a->slen < b->slen);
    *out = (c > 0) - (c < 0);
    return 0;
  }
  if (a->kind == ORC_V_BOOL && b->kind == ORC_V_BOOL) {
    *out = (a->b > b->b) - (a->b < b->b);
    return 0;
  }
  return -1;
}

static int arith(int op, const val *a, const val *b, val *out) {
  memset(out, 0, sizeof *out);
  if (a->kind == ORC_V_NULL || b->kind == ORC_V_NULL) { out->kind = ORC_V_NULL; return 0; }
  if (a->kind == ORC_V_F64 || b->kind == ORC_V_F64) {
    if ((a->kind != ORC_V_F64 && a->kind != ORC_V_I64) || (b->kind != ORC_V_F64 && b->kind != ORC_V_I64)) return -1;
    /* volatile operands: one IEEE operation per node, never contracted into an FMA */
    volatile double x = a->kind == ORC_V_F64 ? a->f : (double)(int64_t)a->i;
    volatile double y = b->kind == ORC_V_F64 ? b->f : (double)(int64_t)b->i;
    volatile double r = op == ORC_X_ADD ? x + y : op == ORC_X_SUB ? x - y : x * y;
    out->kind = ORC_V_F64;
    out->f = r;
    return 0;
  }
  if (a->kind == ORC_V_I128 || b->kind == ORC_V_I128) {
    if ((a->kind != ORC_V_I128 && a->kind != ORC_V_I64) || (b->kind != ORC_V_I128 && b->kind != ORC_V_I64)) return -1;
    /* Decimal128: plain wrapping i128 arithmetic on the unscaled values; the caller is
     * responsible for having rescaled literals (SURVEY 8c decimal type rules) */
    const u128 x = (u128)a->i, y = (u128)b->i;
    out->kind = ORC_V_I128;
    out->i = (i128)(op == ORC_X_ADD ? x + y : op == ORC_X_SUB ? x - y : x * y);
    return 0;
  }
  if (a->kind == ORC_V_I64 && b->kind == ORC_V_I64) {
    const int w = a->width > b->width ? a->width : b->width;
    const u128 x = (u128)a->i, y = (u128)b->i;
    out->kind = ORC_V_I64;
    out->width = w;
    out->i = wrap_int((i128)(op == ORC_X_ADD ? x + y : op == ORC_X_SUB ? x - y : x * y), w);
    return 0;
  }
  return -1;
}

#define STACK_MAX 32
static int eval_expr(const orc_xnode *nodes, int off, int len, const rowctx *ctx, val *out) {
  val st[STACK_MAX];
  int sp = 0;
  for (int k = off; k < off + len; ++k) {
    const orc_xnode *n = &nodes[k];
    if (n->op < ORC_X_ADD) {
      if (sp >= STACK_MAX) return -1;
      val *v = &st[sp++];
      memset(v, 0, sizeof *v);
      switch (n->op) {
        case ORC_X_COL: if (load_col(ctx, n->b, n->a, v)) return -1; break;
        case ORC_X_LIT_F64: v->kind = ORC_V_F64; v->f = n->f; break;
        case ORC_X_LIT_I64: v->kind = ORC_V_I64; v->width = 64; v->i = n->i; break;
        case ORC_X_LIT_I128: v->kind = ORC_V_I128; v->i = (i128)(((u128)(uint64_t)n->i2 << 64) | (uint64_t)n->i); break;
        case ORC_X_LIT_STR: v->kind = ORC_V_STR; v->s = (const uint8_t *)n->s; v->slen = (uint32_t)n->slen; break;
        default: return -1;
      }
      continue;
    }
    if (sp < 2) return -1;
    val b = st[--sp], a = st[--sp], r;
    memset(&r, 0, sizeof r);
    if (n->op <= ORC_X_MUL) {
      if (arith(n->op, &a, &b, &r)) return -1;
    } else if (n->op <= ORC_X_NE) {
      if (a.kind == ORC_V_NULL || b.kind == ORC_V_NULL) {
        r.kind = ORC_V_NULL;
      } else {
        int c;
        if (cmp_vals(&a, &b, &c)) return -1;
        r.kind = ORC_V_BOOL;
        switch (n->op) {
          case ORC_X_LT: r.b = c < 0; break;
          case ORC_X_LE: r.b = c <= 0; break;
          case ORC_X_GT: r.b = c > 0; break;
          case ORC_X_GE: r.b = c >= 0; break;
          case ORC_X_EQ: r.b = c == 0; break;
          default: r.b = c != 0; break;
        }
      }
    } else if (n->op == ORC_X_AND) {
      /* Kleene AND: FALSE dominates NULL */
      const int af = a.kind == ORC_V_BOOL && !a.b, bf = b.kind == ORC_V_BOOL && !b.b;
      if (af || bf) { r.kind = ORC_V_BOOL; r.b = 0; }
      else if (a.kind == ORC_V_NULL || b.kind == ORC_V_NULL) r.kind = ORC_V_NULL;
      else if (a.kind == ORC_V_BOOL && b.kind == ORC_V_BOOL) { r.kind = ORC_V_BOOL; r.b = 1; }
      else return -1;
    } else {
      return -1;
    }
    st[sp++] = r;
  }
  if (sp != 1) return -1;
  *out = st[0];
  return 0;
}

/* FilterExec [DF-K]: keep iff predicate is TRUE (NULL and FALSE dropped) */
static int passes(const orc_xnode *nodes, int off, int len, const rowctx *ctx, int *keep) {
  if (len == 0) { *keep = 1; return 0; }
  val v;
  if (eval_expr(nodes, off, len, ctx, &v)) return -1;
  *keep = v.kind == ORC_V_BOOL && v.b;
  return 0;
}

int orc_filter(const orc_table *scan, const orc_xnode *nodes, int32_t filter_off,
               int32_t filter_len, uint8_t *keep, uint64_t *kept) {
  rowctx ctx;
  memset(&ctx, 0, sizeof ctx);
  ctx.src[0] = scan;
  uint64_t n = 0;
  for (uint64_t r = 0; r < scan->rows; ++r) {
    ctx.row[0] = r;
    int k;
    if (passes(nodes, filter_off, filter_len, &ctx, &k)) return -1;
    keep[r] = (uint8_t)k;
    n += (uint64_t)k;
  }
  if (kept) *kept = n;
  return 0;
}

/* ---- hash join index: key -> chain of build rows in build order ---- */
typedef struct {
  uint64_t nbuckets; /* power of two */
  int64_t *head;     /* bucket -> first row (+1), 0 = empty */
  int64_t *next;     /* row -> next row (+1) in the same bucket */
  const orc_column *key;
} join_index;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

static int int_key(const orc_column *c, uint64_t r, int64_t *out) {
  if (c->validity && !c->validity[r]) return 0; /* NULL keys never match */
  switch (c->type_tag) {
    case ORC_T_INT16: *out = ((const int16_t *)c->values)[r]; return 1;
    case ORC_T_INT32: *out = ((const int32_t *)c->values)[r]; return 1;
    case ORC_T_INT64: *out = ((const int64_t *)c->values)[r]; return 1;
    default: return -1;
  }
}

static int join_index_build(join_index *ix, const orc_table *build, int col) {
  memset(ix, 0, sizeof *ix);
  if ((uint32_t)col >= build->ncols) return -1;
  ix->key = &build->cols[col];
  uint64_t nb = 16;
  while (nb < build->rows * 2) nb <<= 1;
  ix->nbuckets = nb;
  ix->head = calloc(nb, sizeof(int64_t));
  ix->next = calloc(build->rows ? build->rows : 1, sizeof(int64_t));
  if (!ix->head || !ix->next) return -1;
  /* insert in reverse so every chain lists rows in ascending build order */
  for (uint64_t r = build->rows; r-- > 0;) {
    int64_t k;
    const int ok = int_key(ix->key, r, &k);
    if (ok < 0) return -1;
    if (!ok) continue;
    const uint64_t b = mix64((uint64_t)k) & (nb - 1);
    ix->next[r] = ix->head[b];
    ix->head[b] = (int64_t)r + 1;
  }
  return 0;
}

static void join_index_free(join_index *ix) {
  free(ix->head);
  free(ix->next);
}

int orc_hash_join_pairs(const orc_table *build, int32_t build_col, const orc_table *probe,
                        int32_t probe_col, uint64_t *build_rows, uint64_t *probe_rows,
                        uint64_t cap, uint64_t *npairs) {
  join_index ix;
  if (join_index_build(&ix, build, build_col)) { join_index_free(&ix); return -1; }
  if ((uint32_t)probe_col >= probe->ncols) { join_index_free(&ix); return -1; }
  const orc_column *pk = &probe->cols[probe_col];
  uint64_t n = 0;
  for (uint64_t r = 0; r < probe->rows; ++r) {
    int64_t k;
    const int ok = int_key(pk, r, &k);
    if (ok < 0) { join_index_free(&ix); return -1; }
    if (!ok) continue;
    for (int64_t e = ix.head[mix64((uint64_t)k) & (ix.nbuckets - 1)]; e; e = ix.next[e - 1]) {
      int64_t bk = 0;
      int_key(ix.key, (uint64_t)e - 1, &bk);
      if (bk != k) continue;
      if (build_rows && n < cap) { build_rows[n] = (uint64_t)e - 1; probe_rows[n] = r; }
      ++n;
    }
  }
  join_index_free(&ix);
  *npairs = n;
  return 0;
}

/* ---- aggregation ---- */
#define KEY_SLOT 24 /* kind(4) slen(4) payload(16) */
#define MAX_KEYS 8
#define MAX_AGGS 32

typedef struct {
  int kind;      /* ORC_V_F64 / ORC_V_I64 / ORC_V_I128 of the running sum, 0 until first value */
  double fsum;
  i128 isum;
  uint64_t count; /* non-null inputs (or rows for COUNT(*)) */
  val minmax;
  int has_minmax;
  /* no-group lane-striped path */
  double lanes[64];
  uint64_t lane_fill;
} acc_t;

typedef struct {
  uint64_t cap, n;
  uint64_t *slots; /* group index + 1 */
  uint8_t *keys;   /* n * nkeys * KEY_SLOT */
  acc_t *accs;     /* n * naggs */
  uint64_t alloc;
  uint32_t nkeys, naggs;
} group_table;

static uint64_t hash_bytes(const uint8_t *p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  return mix64(h);
}

static int gt_init(group_table *g, uint32_t nkeys, uint32_t naggs) {
  memset(g, 0, sizeof *g);
  g->nkeys = nkeys;
  g->naggs = naggs;
  g->cap = 1024;
  g->alloc = 256;
  g->slots = calloc(g->cap, sizeof(uint64_t));
  g->keys = malloc(g->alloc * (nkeys ? nkeys : 1) * KEY_SLOT);
  g->accs = calloc(g->alloc * (naggs ? naggs : 1), sizeof(acc_t));
  return (g->slots && g->keys && g->accs) ? 0 : -1;
}

static void gt_free(group_table *g) {
  free(g->slots);
  free(g->keys);
  free(g->accs);
}

static int gt_lookup(group_table *g, const uint8_t *key, uint64_t *idx) {
  const size_t klen = (size_t)g->nkeys * KEY_SLOT;
  if ((g->n + 1) * 2 > g->cap) {
    const uint64_t ncap = g->cap * 2;
    uint64_t *ns = calloc(ncap, sizeof(uint64_t));
    if (!ns) return -1;
    for (uint64_t i = 0; i < g->n; ++i) {
      uint64_t b = hash_bytes(g->keys + i * klen, klen) & (ncap - 1);
      while (ns[b]) b = (b + 1) & (ncap - 1);
      ns[b] = i + 1;
    }
    free(g->slots);
    g->slots = ns;
    g->cap = ncap;
  }
  uint64_t b = hash_bytes(key, klen) & (g->cap - 1);
  while (g->slots[b]) {
    const uint64_t i = g->slots[b] - 1;
    if (memcmp(g->keys + i * klen, key, klen) == 0) { *idx = i; return 0; }
    b = (b + 1) & (g->cap - 1);
  }
  if (g->n == g->alloc) {
    const uint64_t na = g->alloc * 2;
    uint8_t *nk = realloc(g->keys, na * (g->nkeys ? g->nkeys : 1) * KEY_SLOT);
    if (!nk) return -1;
    g->keys = nk;
    acc_t *nacc = realloc(g->accs, na * (g->naggs ? g->naggs : 1) * sizeof(acc_t));
    if (!nacc) return -1;
    g->accs = nacc;
    memset(g->accs + g->alloc * (g->naggs ? g->naggs : 1), 0,
           (na - g->alloc) * (g->naggs ? g->naggs : 1) * sizeof(acc_t));
    g->alloc = na;
  }
  memcpy(g->keys + g->n * klen, key, klen);
  g->slots[b] = g->n + 1;
  *idx = g->n++;
  return 0;
}

static int key_encode(const val *v, uint8_t *slot) {
  memset(slot, 0, KEY_SLOT);
  int32_t kind = v->kind, slen = 0;
  switch (v->kind) {
    case ORC_V_NULL: break; /* NULL keys form one group */
    case ORC_V_F64: memcpy(slot + 8, &v->f, 8); break;
    case ORC_V_I64: { int64_t x = (int64_t)v->i; memcpy(slot + 8, &x, 8); break; }
    case ORC_V_I128: memcpy(slot + 8, &v->i, 16); break;
    case ORC_V_BOOL: slot[8] = (uint8_t)v->b; break;
    case ORC_V_STR:
      if (v->slen > 16) return -1; /* oracle limit: group keys up to 16 bytes */
      slen = (int32_t)v->slen;
      memcpy(slot + 8, v->s, v->slen);
      break;
    default: return -1;
  }
  memcpy(slot, &kind, 4);
  memcpy(slot + 4, &slen, 4);
  return 0;
}

static void key_decode(const uint8_t *slot, orc_value *out) {
  memset(out, 0, sizeof *out);
  int32_t kind, slen;
  memcpy(&kind, slot, 4);
  memcpy(&slen, slot + 4, 4);
  out->kind = kind;
  out->slen = slen;
  switch (kind) {
    case ORC_V_F64: memcpy(&out->f, slot + 8, 8); break;
    case ORC_V_I64: memcpy(&out->lo, slot + 8, 8); out->hi = out->lo < 0 ? -1 : 0; break;
    case ORC_V_I128: memcpy(&out->lo, slot + 8, 8); memcpy(&out->hi, slot + 16, 8); break;
    case ORC_V_BOOL: out->lo = slot[8]; break;
    case ORC_V_STR: memcpy(out->s, slot + 8, (size_t)slen); break;
    default: break;
  }
}

typedef struct {
  const orc_xnode *nodes;
  const orc_join_edge *joins;
  join_index *ix;
  uint32_t njoins;
  const int32_t *key_off, *key_len;
  uint32_t nkeys;
  const orc_agg_spec *aggs;
  uint32_t naggs;
  int sum_lanes;
  uint64_t batch_rows;
  group_table gt;
  uint64_t rows_joined;
  uint64_t batch_fill; /* rows in the current no-group batch */
} agg_run;

static void lanes_flush(acc_t *a, int lanes) {
  /* arrow `sum` non-null path [DF-K]: fold the striped lanes pairwise, then add the
   * batch partial to the running accumulator */
  if (a->lane_fill == 0) return;
  int len = lanes;
  while (len >= 2) {
    const int mid = len / 2;
    for (int i = 0; i < mid; ++i) {
      volatile double s = a->lanes[i] + a->lanes[i + mid];
      a->lanes[i] = s;
    }
    len = mid;
  }
  volatile double t = a->fsum + a->lanes[0];
  a->fsum = t;
  memset(a->lanes, 0, sizeof a->lanes);
  a->lane_fill = 0;
}

static int accumulate(agg_run *run, acc_t *accs, const rowctx *ctx) {
  for (uint32_t j = 0; j < run->naggs; ++j) {
    const orc_agg_spec *sp = &run->aggs[j];
    acc_t *a = &accs[j];
    if (sp->func == ORC_AGG_COUNT_STAR) { a->count++; continue; }
    val v;
    if (eval_expr(run->nodes, sp->expr_off, sp->expr_len, ctx, &v)) return -1;
    if (v.kind == ORC_V_NULL) continue; /* aggregates skip NULL inputs */
    a->count++;
    if (sp->func == ORC_AGG_COUNT) continue;
    if (sp->func == ORC_AGG_MIN || sp->func == ORC_AGG_MAX) {
      int c = 0;
      if (a->has_minmax && cmp_vals(&v, &a->minmax, &c)) return -1;
      if (!a->has_minmax || (sp->func == ORC_AGG_MIN ? c < 0 : c > 0)) { a->minmax = v; a->has_minmax = 1; }
      continue;
    }
    /* SUM / AVG */
    if (v.kind == ORC_V_F64 || (sp->func == ORC_AGG_AVG && v.kind == ORC_V_I64)) {
      /* AVG over integers is computed on the Float64 cast [DF-K] */
      const double x = v.kind == ORC_V_F64 ? v.f : (double)(int64_t)v.i;
      a->kind = ORC_V_F64;
      if (run->nkeys == 0 && run->sum_lanes > 0) { /* SumAccumulator and AvgAccumulator both call arrow `sum` per batch */
        const int lane = (int)(a->lane_fill % (uint64_t)run->sum_lanes);
        volatile double s = a->lanes[lane] + x;
        a->lanes[lane] = s;
        a->lane_fill++;
      } else {
        volatile double s = a->fsum + x; /* strictly sequential, input row order */
        a->fsum = s;
      }
    } else if (v.kind == ORC_V_I64) {
      a->kind = ORC_V_I64;
      a->isum = (i128)(int64_t)(uint64_t)((u128)a->isum + (u128)v.i); /* Int64 wrapping */
    } else if (v.kind == ORC_V_I128) {
      a->kind = ORC_V_I128;
      a->isum = (i128)((u128)a->isum + (u128)v.i); /* i128 wrapping */
    } else {
      return -1;
    }
  }
  return 0;
}

static int sink_row(agg_run *run, const rowctx *ctx) {
  uint8_t key[MAX_KEYS * KEY_SLOT];
  for (uint32_t k = 0; k < run->nkeys; ++k) {
    val v;
    if (eval_expr(run->nodes, run->key_off[k], run->key_len[k], ctx, &v)) return -1;
    if (key_encode(&v, key + k * KEY_SLOT)) return -1;
  }
  uint64_t g;
  if (gt_lookup(&run->gt, key, &g)) return -1;
  run->rows_joined++;
  acc_t *accs = run->gt.accs + g * (run->naggs ? run->naggs : 1);
  if (accumulate(run, accs, ctx)) return -1;
  if (run->nkeys == 0 && run->sum_lanes > 0) {
    if (++run->batch_fill == run->batch_rows) {
      for (uint32_t j = 0; j < run->naggs; ++j) lanes_flush(&accs[j], run->sum_lanes);
      run->batch_fill = 0;
    }
  }
  return 0;
}

static int join_level(agg_run *run, rowctx *ctx, uint32_t level) {
  if (level == run->njoins) return sink_row(run, ctx);
  const orc_join_edge *e = &run->joins[level];
  const orc_table *ps = ctx->src[e->probe_src];
  if (!ps || (uint32_t)e->probe_col >= ps->ncols) return -1;
  int64_t k;
  const int ok = int_key(&ps->cols[e->probe_col], ctx->row[e->probe_src], &k);
  if (ok < 0) return -1;
  if (!ok) return 0;
  join_index *ix = &run->ix[level];
  for (int64_t en = ix->head[mix64((uint64_t)k) & (ix->nbuckets - 1)]; en; en = ix->next[en - 1]) {
    int64_t bk = 0;
    int_key(ix->key, (uint64_t)en - 1, &bk);
    if (bk != k) continue;
    ctx->row[level + 1] = (uint64_t)en - 1;
    if (join_level(run, ctx, level + 1)) return -1;
  }
  return 0;
}

int orc_aggregate(const orc_table *scan, const orc_xnode *nodes, int32_t filter_off,
                  int32_t filter_len, const orc_join_edge *joins, uint32_t njoins,
                  const int32_t *key_off, const int32_t *key_len, uint32_t nkeys,
                  const orc_agg_spec *aggs, uint32_t naggs, int32_t sum_lanes,
                  int32_t batch_rows, orc_agg_result *out) {
  memset(out, 0, sizeof *out);
  if (nkeys > MAX_KEYS || naggs > MAX_AGGS || njoins > 7 || sum_lanes > 64) return -1;
  agg_run run;
  memset(&run, 0, sizeof run);
  run.nodes = nodes;
  run.joins = joins;
  run.njoins = njoins;
  run.key_off = key_off;
  run.key_len = key_len;
  run.nkeys = nkeys;
  run.aggs = aggs;
  run.naggs = naggs;
  run.sum_lanes = sum_lanes;
  run.batch_rows = batch_rows > 0 ? (uint64_t)batch_rows : 8192;
  join_index ix[8];
  memset(ix, 0, sizeof ix);
  run.ix = ix;
  int rc = -1;
  if (gt_init(&run.gt, nkeys, naggs)) goto done;
  for (uint32_t j = 0; j < njoins; ++j)
    if (join_index_build(&ix[j], joins[j].build, joins[j].build_col)) goto done;

  rowctx ctx;
  memset(&ctx, 0, sizeof ctx);
  ctx.src[0] = scan;
  for (uint32_t j = 0; j < njoins; ++j) ctx.src[j + 1] = joins[j].build;
  if (nkeys == 0) {
    /* AggregateExec without GROUP BY always emits exactly one row */
    uint64_t g;
    uint8_t dummy = 0;
    if (gt_lookup(&run.gt, &dummy, &g)) goto done;
  }
  for (uint64_t r = 0; r < scan->rows; ++r) {
    ctx.row[0] = r;
    int keep;
    if (passes(nodes, filter_off, filter_len, &ctx, &keep)) goto done;
    if (!keep) continue;
    out->rows_filtered++;
    if (join_level(&run, &ctx, 0)) goto done;
  }
  if (nkeys == 0 && sum_lanes > 0)
    for (uint32_t j = 0; j < naggs; ++j) lanes_flush(&run.gt.accs[j], sum_lanes);

  out->rows_in = scan->rows;
  out->rows_joined = run.rows_joined;
  out->ngroups = run.gt.n;
  out->nkeys = nkeys;
  out->naggs = naggs;
  out->keys = calloc(run.gt.n * (nkeys ? nkeys : 1), sizeof(orc_value));
  out->aggs = calloc(run.gt.n * (naggs ? naggs : 1), sizeof(orc_value));
  if (!out->keys || !out->aggs) goto done;
  for (uint64_t g = 0; g < run.gt.n; ++g) {
    for (uint32_t k = 0; k < nkeys; ++k)
      key_decode(run.gt.keys + (g * nkeys + k) * KEY_SLOT, &out->keys[g * nkeys + k]);
    for (uint32_t j = 0; j < naggs; ++j) {
      const acc_t *a = &run.gt.accs[g * naggs + j];
      orc_value *o = &out->aggs[g * naggs + j];
      switch (aggs[j].func) {
        case ORC_AGG_COUNT_STAR: case ORC_AGG_COUNT:
          o->kind = ORC_V_I64; o->lo = (int64_t)a->count; break;
        case ORC_AGG_SUM:
          if (a->count == 0) { o->kind = ORC_V_NULL; break; } /* SUM of no rows is NULL */
          o->kind = a->kind;
          if (a->kind == ORC_V_F64) o->f = a->fsum;
          else { o->lo = (int64_t)(uint64_t)(u128)a->isum; o->hi = (int64_t)(uint64_t)((u128)a->isum >> 64); }
          break;
        case ORC_AGG_AVG:
          if (a->count == 0) { o->kind = ORC_V_NULL; break; }
          if (a->kind != ORC_V_F64) goto done; /* Decimal AVG: compose from SUM and COUNT */
          o->kind = ORC_V_F64;
          o->f = a->fsum / (double)a->count; /* f64 sum / (u64 count as f64) */
          break;
        case ORC_AGG_MIN: case ORC_AGG_MAX:
          if (!a->has_minmax) { o->kind = ORC_V_NULL; break; }
          o->kind = a->minmax.kind;
          o->f = a->minmax.f;
          o->lo = (int64_t)(uint64_t)(u128)a->minmax.i;
          o->hi = (int64_t)(uint64_t)((u128)a->minmax.i >> 64);
          if (a->minmax.kind == ORC_V_STR) {
            if (a->minmax.slen > 16) goto done;
            o->slen = (int32_t)a->minmax.slen;
            memcpy(o->s, a->minmax.s, a->minmax.slen);
          }
          break;
        default: goto done;
      }
    }
  }
  rc = 0;
done:
  for (uint32_t j = 0; j < njoins; ++j) join_index_free(&ix[j]);
  gt_free(&run.gt);
  if (rc) orc_agg_result_free(out);
  return rc;
}

void orc_agg_result_free(orc_agg_result *r) {
  free(r->keys);
  free(r->aggs);
  memset(r, 0, sizeof *r);
}
