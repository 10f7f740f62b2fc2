// The physical plans DataFusion produces for the TPC-H Q6 / Q1 / Q3 shapes over the reference
// harness schema (benches/tpch/queries/q06.sql, q01.sql, q03.sql; "F" schema of SURVEY 8d: money
// Float64, dates and flags Utf8View), written with the node types of include/pgf_b200_plan.hpp.
// Shared by the CPU lowering test (plan_dump.cpp) and the GPU driver (driver.cpp).
#pragma once
#include "pgf_b200_plan.hpp"

namespace plans {
using namespace pgf_b200;

inline Schema schema_of(std::initializer_list<std::pair<const char*, int32_t>> cols) {
  Schema s;
  for (const auto& c : cols) s.fields.push_back(Field{c.first, c.second, false});  // NOT NULL columns (pg/df_catalog/src/lib.rs:231)
  return s;
}
inline Schema lineitem_q6() {
  return schema_of({{"l_quantity", PGF_T_FLOAT64}, {"l_extendedprice", PGF_T_FLOAT64}, {"l_discount", PGF_T_FLOAT64}, {"l_shipdate", PGF_T_UTF8VIEW}});
}
inline Schema lineitem_q1() {
  return schema_of({{"l_quantity", PGF_T_FLOAT64}, {"l_extendedprice", PGF_T_FLOAT64}, {"l_discount", PGF_T_FLOAT64}, {"l_tax", PGF_T_FLOAT64},
                    {"l_returnflag", PGF_T_UTF8VIEW}, {"l_linestatus", PGF_T_UTF8VIEW}, {"l_shipdate", PGF_T_UTF8VIEW}});
}
inline Schema lineitem_q3() {
  return schema_of({{"l_orderkey", PGF_T_INT32}, {"l_extendedprice", PGF_T_FLOAT64}, {"l_discount", PGF_T_FLOAT64}, {"l_shipdate", PGF_T_UTF8VIEW}});
}
inline Schema orders_q3() {
  return schema_of({{"o_orderkey", PGF_T_INT32}, {"o_custkey", PGF_T_INT32}, {"o_orderdate", PGF_T_UTF8VIEW}, {"o_shippriority", PGF_T_INT32}});
}
inline Schema customer_q3() { return schema_of({{"c_custkey", PGF_T_INT32}, {"c_mktsegment", PGF_T_UTF8VIEW}}); }

inline Schema flags_table() {
  return schema_of({{"active", PGF_T_BOOLEAN}, {"k", PGF_T_INT64}, {"deleted", PGF_T_BOOLEAN}, {"v", PGF_T_FLOAT64}});
}

inline PlanRef scan(uint64_t id, Schema s) { return std::make_shared<WorkerPgScanExec>(id, std::move(s)); }
inline PlanRef filter(ExprRef pred, PlanRef in) {
  return std::make_shared<CoalesceBatchesExec>(std::make_shared<FilterExec>(std::move(pred), std::move(in)));
}

// select sum(l_extendedprice * l_discount) as revenue [, count(*)] from lineitem where l_shipdate >= '1994-01-01'
//   and l_shipdate < '1995-01-01' and l_discount between 0.05 and 0.07 and l_quantity < 24
inline PlanRef q6(uint64_t scan_id) {
  PlanRef li = scan(scan_id, lineitem_q6());
  ExprRef pred = and_(and_(and_(and_(binary(col("l_shipdate", 3), Operator::GtEq, lit("1994-01-01")),
                                     binary(col("l_shipdate", 3), Operator::Lt, lit("1995-01-01"))),
                                binary(col("l_discount", 2), Operator::GtEq, lit(0.05))),
                           binary(col("l_discount", 2), Operator::LtEq, lit(0.07))),
                      binary(col("l_quantity", 0), Operator::Lt, lit(24.0)));
  return std::make_shared<AggregateExec>(
      AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
      std::vector<AggregateFunctionExpr>{sum(binary(col("l_extendedprice", 1), Operator::Multiply, col("l_discount", 2)), "revenue"),
                                         count_star("count(*)")},
      filter(pred, li));
}

// select k, sum(v), count(*) from flags where active and deleted = false and k < 2500 group by k
// (a Boolean column as a predicate by itself, and one compared with a Boolean literal)
inline PlanRef flags_filter(uint64_t scan_id) {
  PlanRef t = scan(scan_id, flags_table());
  ExprRef pred = and_(and_(col("active", 0), binary(col("deleted", 2), Operator::Eq, lit_bool(false))), binary(col("k", 1), Operator::Lt, lit(int64_t(2500))));
  return std::make_shared<AggregateExec>(
      AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{{col("k", 1), "k"}},
      std::vector<AggregateFunctionExpr>{sum(col("v", 3), "sum(v)"), count_star("count(*)")}, filter(pred, t));
}

// q01.sql with the standard eight aggregates.  DataFusion's common-subexpression elimination puts
// l_extendedprice * (1 - l_discount) into a projection below the aggregate; the final projection
// and the sort sit above it.
inline PlanRef q1(uint64_t scan_id, bool partial_final = false) {
  PlanRef li = scan(scan_id, lineitem_q1());
  PlanRef filtered = filter(binary(col("l_shipdate", 6), Operator::LtEq, lit("1998-09-02")), li);
  ExprRef disc_price = binary(col("l_extendedprice", 1), Operator::Multiply, binary(lit(1.0), Operator::Minus, col("l_discount", 2)));
  PlanRef proj = std::make_shared<ProjectionExec>(
      std::vector<std::pair<ExprRef, std::string>>{{disc_price, "__common_expr_1"}, {col("l_quantity", 0), "l_quantity"},
                                                   {col("l_extendedprice", 1), "l_extendedprice"}, {col("l_discount", 2), "l_discount"},
                                                   {col("l_tax", 3), "l_tax"}, {col("l_returnflag", 4), "l_returnflag"},
                                                   {col("l_linestatus", 5), "l_linestatus"}},
      filtered);
  std::vector<std::pair<ExprRef, std::string>> gby{{col("l_returnflag", 5), "l_returnflag"}, {col("l_linestatus", 6), "l_linestatus"}};
  std::vector<AggregateFunctionExpr> aggs{
      sum(col("l_quantity", 1), "sum_qty"), sum(col("l_extendedprice", 2), "sum_base_price"), sum(col("__common_expr_1", 0), "sum_disc_price"),
      sum(binary(col("__common_expr_1", 0), Operator::Multiply, binary(lit(1.0), Operator::Plus, col("l_tax", 4))), "sum_charge"),
      avg(col("l_quantity", 1), "avg_qty"), avg(col("l_extendedprice", 2), "avg_price"), avg(col("l_discount", 3), "avg_disc"),
      count_star("count_order")};
  PlanRef agg;
  if (partial_final) {  // the shape DataFusion plans with more than one partition, collapsed over one
    PlanRef partial = std::make_shared<AggregateExec>(AggregateMode::Partial, gby, aggs, proj);
    agg = std::make_shared<AggregateExec>(AggregateMode::FinalPartitioned, gby, aggs, std::make_shared<CoalesceBatchesExec>(partial));
  } else {
    agg = std::make_shared<AggregateExec>(AggregateMode::Single, gby, aggs, proj);
  }
  std::vector<PhysicalSortExpr> order{sort_asc(col("l_returnflag", 0)), sort_asc(col("l_linestatus", 1))};
  return std::make_shared<SortExec>(order, agg);
}

// ---- "D" variants: money Decimal128(15,2), dates Date32 (Int32 days), flags Int16 codes.
// Literals are what DataFusion's type coercion leaves in the physical plan: unscaled Decimal128
// values at the column's scale (0.05 -> 5, 24 -> 2400, 1 -> 100) and Date32 day numbers.
inline Schema lineitem_q6_d() {
  return schema_of({{"l_quantity", PGF_T_DECIMAL128}, {"l_extendedprice", PGF_T_DECIMAL128}, {"l_discount", PGF_T_DECIMAL128}, {"l_shipdate", PGF_T_INT32}});
}
inline Schema lineitem_q1_d() {
  return schema_of({{"l_quantity", PGF_T_DECIMAL128}, {"l_extendedprice", PGF_T_DECIMAL128}, {"l_discount", PGF_T_DECIMAL128}, {"l_tax", PGF_T_DECIMAL128},
                    {"l_returnflag", PGF_T_INT16}, {"l_linestatus", PGF_T_INT16}, {"l_shipdate", PGF_T_INT32}});
}
inline ExprRef dec(int64_t unscaled) { return lit(ScalarValue::decimal128(unscaled)); }

inline PlanRef q6_d(uint64_t scan_id) {
  PlanRef li = scan(scan_id, lineitem_q6_d());
  ExprRef pred = and_(and_(and_(and_(binary(col("l_shipdate", 3), Operator::GtEq, lit(int64_t(8766))),    // 1994-01-01
                                     binary(col("l_shipdate", 3), Operator::Lt, lit(int64_t(9131)))),     // 1995-01-01
                                binary(col("l_discount", 2), Operator::GtEq, dec(5))),
                           binary(col("l_discount", 2), Operator::LtEq, dec(7))),
                      binary(col("l_quantity", 0), Operator::Lt, dec(2400)));
  return std::make_shared<AggregateExec>(
      AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
      std::vector<AggregateFunctionExpr>{sum(binary(col("l_extendedprice", 1), Operator::Multiply, col("l_discount", 2)), "revenue"),
                                         count_star("count(*)")},
      filter(pred, li));
}

inline PlanRef q1_d(uint64_t scan_id) {
  PlanRef li = scan(scan_id, lineitem_q1_d());
  PlanRef filtered = filter(binary(lit(int64_t(10471)), Operator::GtEq, col("l_shipdate", 6)), li);   // 1998-09-02 >= l_shipdate (flipped on purpose)
  ExprRef disc_price = binary(col("l_extendedprice", 1), Operator::Multiply, binary(dec(100), Operator::Minus, col("l_discount", 2)));
  std::vector<AggregateFunctionExpr> aggs{
      sum(col("l_quantity", 0), "sum_qty"), sum(col("l_extendedprice", 1), "sum_base_price"), sum(disc_price, "sum_disc_price"),
      sum(binary(disc_price, Operator::Multiply, binary(col("l_tax", 3), Operator::Plus, dec(100))), "sum_charge"),   // (x + c) form
      avg(col("l_quantity", 0), "avg_qty"), avg(col("l_extendedprice", 1), "avg_price"), avg(col("l_discount", 2), "avg_disc"),
      count_star("count_order")};
  return std::make_shared<AggregateExec>(
      AggregateMode::Single,
      std::vector<std::pair<ExprRef, std::string>>{{col("l_returnflag", 4), "l_returnflag"}, {col("l_linestatus", 5), "l_linestatus"}}, aggs, filtered);
}

// q03.sql: customer(BUILDING) |><| orders(o_orderdate < d) |><| lineitem(l_shipdate > d),
// group by l_orderkey, o_orderdate, o_shippriority order by revenue desc, o_orderdate limit 10
inline PlanRef q3(uint64_t customer_id, uint64_t orders_id, uint64_t lineitem_id, uint64_t fetch = 10, const char* segment = "BUILDING") {
  PlanRef c = filter(binary(col("c_mktsegment", 1), Operator::Eq, lit(segment)), scan(customer_id, customer_q3()));
  PlanRef o = filter(binary(col("o_orderdate", 2), Operator::Lt, lit("1995-03-15")), scan(orders_id, orders_q3()));
  PlanRef l = filter(binary(col("l_shipdate", 3), Operator::Gt, lit("1995-03-15")), scan(lineitem_id, lineitem_q3()));
  // customer ++ orders = c_custkey, c_mktsegment, o_orderkey, o_custkey, o_orderdate, o_shippriority -> keep 2, 4, 5
  PlanRef co = std::make_shared<HashJoinExec>(c, o, HashJoinExec::JoinOn{{col("c_custkey", 0), col("o_custkey", 1)}}, JoinType::Inner,
                                              PartitionMode::CollectLeft, false, nullptr, std::vector<size_t>{2, 4, 5});
  // (o_orderkey, o_orderdate, o_shippriority) ++ lineitem
  PlanRef col_ = std::make_shared<HashJoinExec>(std::make_shared<CoalesceBatchesExec>(co), l,
                                                HashJoinExec::JoinOn{{col("o_orderkey", 0), col("l_orderkey", 0)}});
  PlanRef agg = std::make_shared<AggregateExec>(
      AggregateMode::Single,
      std::vector<std::pair<ExprRef, std::string>>{{col("l_orderkey", 3), "l_orderkey"}, {col("o_orderdate", 1), "o_orderdate"},
                                                   {col("o_shippriority", 2), "o_shippriority"}},
      std::vector<AggregateFunctionExpr>{
          sum(binary(col("l_extendedprice", 4), Operator::Multiply, binary(lit(1.0), Operator::Minus, col("l_discount", 5))), "revenue")},
      std::make_shared<CoalesceBatchesExec>(col_));
  PlanRef proj = std::make_shared<ProjectionExec>(
      std::vector<std::pair<ExprRef, std::string>>{{col("l_orderkey", 0), "l_orderkey"}, {col("revenue", 3), "revenue"},
                                                   {col("o_orderdate", 1), "o_orderdate"}, {col("o_shippriority", 2), "o_shippriority"}},
      agg);
  std::vector<PhysicalSortExpr> order{sort_desc(col("revenue", 1)), sort_asc(col("o_orderdate", 2))};
  if (fetch == 0) return std::make_shared<SortExec>(order, proj);
  return std::make_shared<SortExec>(order, proj, fetch);
}

// The same three-way join planned right-deep: customer |><| (orders |><| lineitem) -- the lineitem stream probes the
// orders table and, for the rows that matched, the customer table with the matched order's o_custkey.  Two probes fused
// into one scan stream; the same joined rows as q3().
inline PlanRef q3_right_deep_count(uint64_t customer_id, uint64_t orders_id, uint64_t lineitem_id, const char* segment = "BUILDING") {
  PlanRef c = filter(binary(col("c_mktsegment", 1), Operator::Eq, lit(segment)), scan(customer_id, customer_q3()));
  PlanRef o = filter(binary(col("o_orderdate", 2), Operator::Lt, lit("1995-03-15")), scan(orders_id, orders_q3()));
  PlanRef l = filter(binary(col("l_shipdate", 3), Operator::Gt, lit("1995-03-15")), scan(lineitem_id, lineitem_q3()));
  PlanRef inner = std::make_shared<HashJoinExec>(o, l, HashJoinExec::JoinOn{{col("o_orderkey", 0), col("l_orderkey", 0)}});
  PlanRef outer = std::make_shared<HashJoinExec>(c, inner, HashJoinExec::JoinOn{{col("c_custkey", 0), col("o_custkey", 1)}});
  return std::make_shared<AggregateExec>(AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
                                         std::vector<AggregateFunctionExpr>{count_star("joined_rows")}, outer);
}

}  // namespace plans
