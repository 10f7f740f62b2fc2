// CPU test of the host-side planner (include/pgf_b200_plan.hpp): builds the Q6 / Q1 / Q3 physical
// plans, runs install_runtime_filters + install_b200_operators without a device and prints, for
// tests/test_cpp_host.py to check:
//   TREE <name> / indented plan / END        the rewritten plan (DisplayAs)
//   POD <name> <index> <hex of pgf_pipeline>  every lowered pipeline, build sides first
//   CHECK <name> ok|FAILED                    host-side behaviours (errors, eligibility)
#include <cstdio>
#include <iostream>

#include "plans.hpp"

using namespace pgf_b200;

namespace {

int failures = 0;
void check(const char* name, bool ok) {
  std::printf("CHECK %s %s\n", name, ok ? "ok" : "FAILED");
  if (!ok) ++failures;
}

// A pool that hands out fixed handles, so the lowering can be compared byte for byte.
class FakePool final : public RuntimeFilterPool {
 public:
  explicit FakePool(uint32_t slots) : slots_(slots) {}
  std::optional<RuntimeFilterBuildHandle> allocate_build(const RuntimeFilterTarget& t) override {
    if (targets.size() >= slots_) return std::nullopt;
    targets.push_back(t);
    return RuntimeFilterBuildHandle{100 + targets.size(), 1};
  }
  std::vector<RuntimeFilterTarget> targets;

 private:
  uint32_t slots_;
};

void dump(const char* name, const PlanRef& plan) {
  std::printf("TREE %s\n%sEND\n", name, display_indent(plan).c_str());
  std::vector<const B200PipelineExec*> pods;
  detail::collect_pods(plan, pods);
  for (size_t i = 0; i < pods.size(); ++i) {
    std::printf("POD %s %zu ", name, i);
    const auto* bytes = reinterpret_cast<const unsigned char*>(&pods[i]->pod());
    for (size_t b = 0; b < sizeof(pgf_pipeline); ++b) std::printf("%02x", bytes[b]);
    std::printf("\n");
  }
}

template <class F>
bool throws(ErrorKind kind, F&& f) {
  try {
    f();
  } catch (const DataFusionError& e) {
    return e.kind() == kind;
  } catch (...) {
    return false;
  }
  return false;
}

}  // namespace

int main() {
  std::vector<std::string> skipped;
  dump("q6", install_b200_operators(plans::q6(1), nullptr, &skipped));
  dump("q1", install_b200_operators(plans::q1(2), nullptr, &skipped));
  dump("q1_partial_final", install_b200_operators(plans::q1(2, true), nullptr, &skipped));
  dump("q3", install_b200_operators(plans::q3(3, 4, 5), nullptr, &skipped));
  dump("q6_decimal", install_b200_operators(plans::q6_d(6), nullptr, &skipped));
  dump("q1_decimal", install_b200_operators(plans::q1_d(7), nullptr, &skipped));
  dump("flags_filter", install_b200_operators(plans::flags_filter(8), nullptr, &skipped));
  {
    FakePool pool(64);
    PlanRef with_filters = install_runtime_filters(plans::q3(3, 4, 5), 7, pool);
    std::printf("TREE q3_filters_before\n%sEND\n", display_indent(with_filters).c_str());
    dump("q3_filters", install_b200_operators(with_filters, nullptr, &skipped));
    // the inner join is visited first; targets name the probe-side scan and its key column
    check("runtime_filter_targets", pool.targets.size() == 2 && pool.targets[0].scan_id == 4 && pool.targets[0].output_column == 1 &&
                                        pool.targets[0].key_type == 2 && pool.targets[0].session_epoch == 7 &&
                                        pool.targets[1].scan_id == 5 && pool.targets[1].output_column == 0);
  }
  {
    FakePool pool(1);  // exhausted after the first join: a soft miss, the plan is still valid
    PlanRef p = install_b200_operators(install_runtime_filters(plans::q3(3, 4, 5), 7, pool), nullptr, &skipped);
    dump("q3_pool_exhausted", p);
    check("pool_exhaustion_is_soft", pool.targets.size() == 1 && p->downcast<B200PipelineExec>() != nullptr);
  }
  check("everything_absorbed", skipped.empty());

  // -- outside the grammar: the DataFusion node stays, with the reason recorded
  {
    PlanRef li = plans::scan(1, plans::lineitem_q6());
    ExprRef disj = binary(binary(col("l_discount", 2), Operator::Lt, lit(0.05)), Operator::Or, binary(col("l_quantity", 0), Operator::Lt, lit(24.0)));
    PlanRef agg = std::make_shared<AggregateExec>(AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
                                                  std::vector<AggregateFunctionExpr>{count_star("n")}, plans::filter(disj, li));
    std::vector<std::string> why;
    PlanRef out = install_b200_operators(agg, nullptr, &why);
    check("or_predicate_not_absorbed", out->downcast<AggregateExec>() != nullptr && why.size() == 1);
    TaskContext tc(nullptr);
    check("unabsorbed_node_has_no_cpu_operator", throws(ErrorKind::NotImplemented, [&] { out->execute(0, tc); }));
  }
  {
    PlanRef c = plans::scan(3, plans::customer_q3()), o = plans::scan(4, plans::orders_q3());
    PlanRef left_join = std::make_shared<HashJoinExec>(c, o, HashJoinExec::JoinOn{{col("c_custkey", 0), col("o_custkey", 1)}}, JoinType::Left);
    PlanRef agg = std::make_shared<AggregateExec>(AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
                                                  std::vector<AggregateFunctionExpr>{count_star("n")}, left_join);
    FakePool pool(4);
    std::vector<std::string> why;
    PlanRef out = install_b200_operators(install_runtime_filters(agg, 1, pool), nullptr, &why);
    check("outer_join_gets_no_filter_and_is_not_absorbed", pool.targets.empty() && out->downcast<AggregateExec>() != nullptr && why.size() == 1);
    // a string join key is not eligible for a runtime filter (key_type_for), nor for the fused join
    PlanRef sj = std::make_shared<HashJoinExec>(c, c, HashJoinExec::JoinOn{{col("c_mktsegment", 1), col("c_mktsegment", 1)}});
    install_runtime_filters(sj, 1, pool);
    check("string_key_gets_no_filter", pool.targets.empty());
    // null_equals_null joins are skipped like in maybe_wrap_hash_join
    PlanRef nn = std::make_shared<HashJoinExec>(c, o, HashJoinExec::JoinOn{{col("c_custkey", 0), col("o_custkey", 1)}}, JoinType::Inner,
                                                PartitionMode::CollectLeft, true);
    install_runtime_filters(nn, 1, pool);
    check("null_equals_null_gets_no_filter", pool.targets.empty());
  }
  {
    // right-deep chain: lineitem probes orders and then customer-like table on one stream -> two probes
    PlanRef c = plans::scan(3, plans::customer_q3()), o = plans::scan(4, plans::orders_q3()), l = plans::scan(5, plans::lineitem_q3());
    PlanRef inner = std::make_shared<HashJoinExec>(o, l, HashJoinExec::JoinOn{{col("o_orderkey", 0), col("l_orderkey", 0)}});
    PlanRef outer = std::make_shared<HashJoinExec>(c, inner, HashJoinExec::JoinOn{{col("c_custkey", 0), col("o_custkey", 1)}});
    PlanRef agg = std::make_shared<AggregateExec>(AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
                                                  std::vector<AggregateFunctionExpr>{count_star("n")}, outer);
    std::vector<std::string> why;
    PlanRef out = install_b200_operators(agg, nullptr, &why);
    check("two_probes_on_one_stream_are_fused", out->downcast<B200PipelineExec>() != nullptr && why.empty());
    dump("two_probes", out);
    // a third probe on the same stream is one too many: the plan keeps its DataFusion nodes
    PlanRef c2 = plans::scan(6, plans::customer_q3());
    PlanRef third = std::make_shared<HashJoinExec>(c2, outer, HashJoinExec::JoinOn{{col("c_custkey", 0), col("c_custkey", 0)}});
    PlanRef agg3 = std::make_shared<AggregateExec>(AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
                                                   std::vector<AggregateFunctionExpr>{count_star("n")}, third);
    why.clear();
    PlanRef out3 = install_b200_operators(agg3, nullptr, &why);
    check("three_probes_on_one_stream_stay_datafusion", out3->downcast<AggregateExec>() != nullptr && !why.empty());
  }
  {
    // limits of the fused kernels that the grammar mirrors, so that joined pipelines (which the library
    // can only check once their tables exist) are decided at plan time too
    auto count_over = [](PlanRef in) {
      return PlanRef(std::make_shared<AggregateExec>(AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
                                                     std::vector<AggregateFunctionExpr>{count_star("n")}, std::move(in)));
    };
    std::vector<std::string> why;
    PlanRef li = plans::scan(1, plans::lineitem_q6());
    PlanRef long_lit = count_over(plans::filter(binary(col("l_shipdate", 3), Operator::Lt, lit("1995-01-01 00:00")), li));
    check("long_string_literal_not_absorbed", install_b200_operators(long_lit, nullptr, &why)->downcast<AggregateExec>() != nullptr);
    // predicate on a build-side column above the join
    PlanRef c = plans::scan(3, plans::customer_q3()), o = plans::scan(4, plans::orders_q3());
    PlanRef join = std::make_shared<HashJoinExec>(c, o, HashJoinExec::JoinOn{{col("c_custkey", 0), col("o_custkey", 1)}});
    PlanRef build_pred = count_over(plans::filter(binary(col("c_mktsegment", 1), Operator::Eq, lit("BUILDING")), join));
    check("predicate_on_build_column_not_absorbed", install_b200_operators(build_pred, nullptr, &why)->downcast<AggregateExec>() != nullptr);
    // build-side columns above the join wider than a slot: two views = 32 bytes > 20
    Schema wide = plans::schema_of({{"k", PGF_T_INT32}, {"a", PGF_T_UTF8VIEW}, {"b", PGF_T_UTF8VIEW}});
    PlanRef wj = std::make_shared<HashJoinExec>(plans::scan(8, wide), o, HashJoinExec::JoinOn{{col("k", 0), col("o_custkey", 3 - 2)}});
    PlanRef wagg = std::make_shared<AggregateExec>(
        AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{{col("a", 1), "a"}, {col("b", 2), "b"}},
        std::vector<AggregateFunctionExpr>{count_star("n")}, wj);
    check("wide_payload_not_absorbed", install_b200_operators(wagg, nullptr, &why)->downcast<AggregateExec>() != nullptr);
    // group key wider than 32 bytes: three views
    Schema three = plans::schema_of({{"a", PGF_T_UTF8VIEW}, {"b", PGF_T_UTF8VIEW}, {"c", PGF_T_UTF8VIEW}});
    PlanRef kagg = std::make_shared<AggregateExec>(
        AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{{col("a", 0), "a"}, {col("b", 1), "b"}, {col("c", 2), "c"}},
        std::vector<AggregateFunctionExpr>{count_star("n")}, plans::scan(9, three));
    check("wide_group_key_not_absorbed", install_b200_operators(kagg, nullptr, &why)->downcast<AggregateExec>() != nullptr);
    // arithmetic on narrow integers wraps at their own width in arrow: not computed here
    PlanRef l3 = plans::scan(5, plans::lineitem_q3());
    PlanRef narrow = std::make_shared<AggregateExec>(
        AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
        std::vector<AggregateFunctionExpr>{sum(binary(col("l_orderkey", 0), Operator::Multiply, col("l_orderkey", 0)), "s")}, l3);
    check("int32_arithmetic_not_absorbed", install_b200_operators(narrow, nullptr, &why)->downcast<AggregateExec>() != nullptr);
    PlanRef mixed = std::make_shared<AggregateExec>(
        AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
        std::vector<AggregateFunctionExpr>{sum(binary(col("l_orderkey", 0), Operator::Multiply, col("l_discount", 2)), "s")}, l3);
    check("mixed_type_arithmetic_not_absorbed", install_b200_operators(mixed, nullptr, &why)->downcast<AggregateExec>() != nullptr);
    PlanRef plain = std::make_shared<AggregateExec>(
        AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{},
        std::vector<AggregateFunctionExpr>{sum(col("l_orderkey", 0), "s")}, l3);
    check("plain_int32_sum_is_absorbed", install_b200_operators(plain, nullptr, &why)->downcast<B200PipelineExec>() != nullptr);
    check("limits_record_their_reasons", why.size() == 6);
  }
  {
    // ResultPageProducer over a hand-built aggregate result: 10 000 rows of (Int64 key, nullable Float64 sum,
    // count) leave as pages of the fixed row cap, then one close step, then nothing
    const uint64_t n = 10000;
    auto* r = new pgf_result();
    r->ngroups = n; r->nkeys = 1; r->naggs = 2;
    r->keys = new pgf_value[n + 1]();
    r->aggs = new pgf_value[2 * n + 1]();
    r->key_type[0] = PGF_T_INT64; r->agg_type[0] = PGF_T_FLOAT64; r->agg_type[1] = PGF_T_INT64;
    r->agg_func[0] = PGF_AGG_SUM; r->agg_func[1] = PGF_AGG_COUNT_STAR;
    for (uint64_t g = 0; g < n; ++g) {
      r->keys[g].kind = PGF_V_I64; r->keys[g].lo = int64_t(g) - 5000;
      if (g % 7 == 3) r->aggs[2 * g].kind = PGF_V_NULL;
      else { r->aggs[2 * g].kind = PGF_V_F64; r->aggs[2 * g].f64 = double(g) * 0.5; }
      r->aggs[2 * g + 1].kind = PGF_V_I64; r->aggs[2 * g + 1].lo = int64_t(g);
    }
    RecordBatch batch;
    batch.num_rows = n;
    batch.raw = std::shared_ptr<pgf_result>(r, [](pgf_result* p) { pgf_result_free(p); });
    ResultPageProducer producer(batch, 65536);
    const auto& schema = producer.transport_schema();
    check("transport_schema", schema.size() == 3 && schema[0].type_tag == PGF_T_INT64 && schema[0].nullable == 1 &&
                                  schema[1].type_tag == PGF_T_FLOAT64 && schema[1].nullable == 1 && schema[2].nullable == 0);
    uint64_t pages = 0, rows = 0;
    bool pages_ok = true, closed = false;
    while (auto step = producer.next_step()) {
      if (step->kind == ResultPageStep::CloseFrame) { closed = true; continue; }
      pages_ok = pages_ok && !closed && step->rows <= producer.rows_per_page();
      uint16_t kind = 0, flags = 0;
      uint32_t len = 0;
      pages_ok = pages_ok && pgf_page_header_decode(step->page.data(), &kind, &flags, &len) == PGF_OK && kind == PGF_ARROW_LAYOUT_BATCH_KIND;
      pages_ok = pages_ok && pgf_block_import_check(kind, flags, step->page.data() + PGF_PAGE_HEADER_LEN, len, schema.data(), 3) == PGF_OK;
      ++pages;
      rows += step->rows;
    }
    const uint64_t cap = producer.rows_per_page();
    check("result_pages_one_per_step", pages_ok && closed && rows == n && pages == (n + cap - 1) / cap && pages > 1);
    check("nothing_after_the_close_step", !producer.next_step() && !producer.next_step());
    RecordBatch empty;
    auto* e = new pgf_result();
    e->nkeys = 1; e->key_type[0] = PGF_T_INT64;
    e->keys = new pgf_value[1]();
    e->aggs = new pgf_value[1]();
    empty.raw = std::shared_ptr<pgf_result>(e, [](pgf_result* p) { pgf_result_free(p); });
    ResultPageProducer none(empty, 65536);
    auto first = none.next_step();
    check("empty_result_goes_straight_to_close", first && first->kind == ResultPageStep::CloseFrame && !none.next_step());
    RecordBatch foreign;
    check("producer_needs_a_pipeline_result", throws(ErrorKind::Execution, [&] { ResultPageProducer p(foreign, 65536); }));
  }
  // -- the node surface
  {
    PlanRef p = install_b200_operators(plans::q6(1), nullptr);
    TaskContext tc(nullptr);
    check("one_partition", p->partition_count() == 1);
    check("partition_1_is_a_plan_error", throws(ErrorKind::Plan, [&] { p->execute(1, tc); }));
    check("no_device_is_an_execution_error", throws(ErrorKind::Execution, [&] { p->execute(0, tc); }));
    check("with_new_children_arity", throws(ErrorKind::Plan, [&] { p->with_new_children({}); }));
    PlanRef same = p->with_new_children(p->children());
    check("with_new_children_roundtrip", same->downcast<B200PipelineExec>() &&
                                             std::memcmp(&same->downcast<B200PipelineExec>()->pod(), &p->downcast<B200PipelineExec>()->pod(), sizeof(pgf_pipeline)) == 0);
    int32_t devices = 0;
    pgf_device_count(&devices);
    if (devices == 0)  // the library refuses to create a context without a device: there is no CPU fallback
      check("no_device_context", throws(ErrorKind::Execution, [] { B200Context ctx(0); (void)ctx; }));
  }
  return failures ? 1 : 0;
}
