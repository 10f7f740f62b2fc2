// GPU driver for the C++ host layer (include/pgf_b200_plan.hpp): generates TPC-H shaped scans in
// HBM, plans the Q6 / Q1 / Q3 physical plans of plans.hpp, rewrites them with
// install_runtime_filters + install_b200_operators and executes them.  One JSON object per line on
// stdout; tests/test_gpu_cpp_host.py compares every value with the CPU oracle run over the same
// generated pages.
//   driver <q6_rows> <q1_rows> <customers> <orders> <lineitems> [result_pages_out]
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>

#include "plans.hpp"

using namespace pgf_b200;

namespace {

void gen(B200Context& gpu, uint64_t scan_id, int32_t table, uint64_t rows, uint64_t scale_rows) {
  pgf_gen_spec spec{table, 0, 42, 0, rows, scale_rows};
  gpu.check(pgf_gen_scan(gpu.raw(), scan_id, &spec));
}

void print_value(const pgf_value& v) {
  switch (v.kind) {
    case PGF_V_NULL: std::printf("null"); break;
    case PGF_V_F64: std::printf("%.17g", v.f64); break;
    case PGF_V_I64: std::printf("%" PRId64, v.lo); break;
    case PGF_V_I128: std::printf("[%" PRId64 ", %" PRIu64 "]", v.hi, uint64_t(v.lo)); break;
    case PGF_V_STR: std::printf("\"%.*s\"", v.slen, reinterpret_cast<const char*>(v.str)); break;
    default: std::printf("\"?\"");
  }
}

void print_run(const char* query, const PlanRef& plan, const RecordBatch& batch, const TaskContext& tc) {
  std::printf("{\"query\": \"%s\", \"root\": \"%s\", \"columns\": [", query, plan->name().c_str());
  for (size_t c = 0; c < batch.schema.size(); ++c) std::printf("%s\"%s\"", c ? ", " : "", batch.schema.fields[c].name.c_str());
  std::printf("], \"rows\": [");
  for (uint64_t r = 0; r < batch.num_rows; ++r) {
    std::printf("%s[", r ? ", " : "");
    for (size_t c = 0; c < batch.columns.size(); ++c) {
      if (c) std::printf(", ");
      print_value(batch.columns[c][r]);
    }
    std::printf("]");
  }
  std::printf("], \"pipelines\": [");
  for (size_t i = 0; i < tc.metrics.size(); ++i) {
    const PipelineMetrics& m = tc.metrics[i];
    std::printf("%s{\"rows_in\": %" PRIu64 ", \"rows_bloom\": %" PRIu64 ", \"rows_filtered\": %" PRIu64 ", \"rows_out\": %" PRIu64
                ", \"bloom_rows\": %" PRIu64 ", \"kernel_launches\": %u, \"variant\": \"%s\"}",
                i ? ", " : "", m.rows_in, m.rows_bloom, m.rows_filtered, m.rows_out, m.bloom_rows, m.kernel_launches, m.variant.c_str());
  }
  std::printf("]}\n");
}

RecordBatch run(const char* query, B200Context& gpu, const PlanRef& physical, RuntimeFilterPool* pool = nullptr) {
  PlanRef plan = physical;
  if (pool) plan = install_runtime_filters(plan, /*session_epoch=*/1, *pool);
  std::vector<std::string> skipped;
  plan = install_b200_operators(plan, &gpu, &skipped);
  for (const auto& s : skipped) std::fprintf(stderr, "%s: not absorbed: %s\n", query, s.c_str());
  TaskContext tc(&gpu);
  RecordBatch batch = plan->execute(0, tc);
  print_run(query, plan, batch, tc);
  return batch;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: driver q6_rows q1_rows customers orders lineitems [result_pages_out]\n");
    return 2;
  }
  const uint64_t q6_rows = std::strtoull(argv[1], nullptr, 10), q1_rows = std::strtoull(argv[2], nullptr, 10);
  const uint64_t ncust = std::strtoull(argv[3], nullptr, 10), nord = std::strtoull(argv[4], nullptr, 10), nli = std::strtoull(argv[5], nullptr, 10);
  try {
    // every fused Bloom probe is kept, so the runtime-filter plan shows its effect in rows_bloom
    B200Context gpu(0, 65536, PGF_CFG_KEEP_REDUNDANT_BLOOM_PROBES);
    gen(gpu, 1, PGF_GEN_LINEITEM_Q6, q6_rows, 0);
    gen(gpu, 2, PGF_GEN_LINEITEM_Q1, q1_rows, 0);
    gen(gpu, 3, PGF_GEN_CUSTOMER_Q3, ncust, 0);
    gen(gpu, 4, PGF_GEN_ORDERS_Q3, nord, ncust);
    gen(gpu, 5, PGF_GEN_LINEITEM_Q3, nli, nord);

    run("q6", gpu, plans::q6(1));
    RecordBatch q1 = run("q1", gpu, plans::q1(2));
    run("q1_partial_final", gpu, plans::q1(2, true));
    run("q3", gpu, plans::q3(3, 4, 5));
    run("q3_all_groups", gpu, plans::q3(3, 4, 5, /*fetch=*/0));
    run("q3_two_probes", gpu, plans::q3_right_deep_count(3, 4, 5));
    {
      pgf_bloom_params params;
      gpu.check(pgf_bloom_params_new(1u << 20, 4, 0x7067667573696f6eull, &params));  // GUC defaults, pg/extension/src/guc.rs:41-46
      DeviceRuntimeFilterPool pool(gpu, params);
      run("q3_runtime_filters", gpu, plans::q3(3, 4, 5), &pool);
    }
    // "D" variants: exact Decimal128 sums
    gen(gpu, 6, PGF_GEN_LINEITEM_Q6_D, q6_rows, 0);
    gen(gpu, 7, PGF_GEN_LINEITEM_Q1_D, q1_rows, 0);
    run("q6_decimal", gpu, plans::q6_d(6));
    run("q1_decimal", gpu, plans::q1_d(7));
    // Q3 again over scans that are fed page by page from one producer thread per scan, the way the
    // worker's scan threads deliver transfer pages: declare -> push_page* (concurrently) -> finish
    {
      struct Feed { uint64_t from, to; Schema schema; std::vector<uint8_t> pages; pgf_scan_info info; };
      Feed feeds[3] = {{3, 13, plans::customer_q3(), {}, {}}, {4, 14, plans::orders_q3(), {}, {}}, {5, 15, plans::lineitem_q3(), {}, {}}};
      for (Feed& f : feeds) {
        gpu.check(pgf_scan_get_info(gpu.raw(), f.from, &f.info));
        f.pages.resize(size_t(f.info.pages) * gpu.page_size());
        gpu.check(pgf_scan_read_pages(gpu.raw(), f.from, 0, f.info.pages, f.pages.data()));
      }
      // scans are declared when the plan is made; every producer thread pushes its pages and ends
      // its own stream (push and finish of different scans may run concurrently)
      std::vector<ScanIngest> ingests;
      for (Feed& f : feeds) ingests.emplace_back(gpu, f.to, f.schema, f.info.pages);
      std::vector<std::thread> producers;
      std::string errors[3];
      for (int i = 0; i < 3; ++i) {
        producers.emplace_back([&, i] {
          try {
            const Feed& f = feeds[i];
            ScanIngest& ingest = ingests[size_t(i)];
            for (uint64_t p = 0; p < f.info.pages; ++p) ingest.push_page(f.pages.data() + p * gpu.page_size(), gpu.page_size());
            ingest.finish();
            if (ingest.info().rows != f.info.rows) errors[i] = "row count differs after ingest";
          } catch (const DataFusionError& e) {
            errors[i] = e.what();
          }
        });
      }
      for (auto& t : producers) t.join();
      for (const auto& e : errors)
        if (!e.empty()) throw exec_err("streamed ingest: " + e);
      run("q3_streamed", gpu, plans::q3(13, 14, 15));
    }
    if (argc > 6) {  // ResultPageProducer: the Q1 rows as transfer pages
      uint64_t npages = 0;
      const std::vector<uint8_t> pages = encode_result_pages(q1, gpu.page_size(), &npages);
      FILE* f = std::fopen(argv[6], "wb");
      if (!f || std::fwrite(pages.data(), 1, pages.size(), f) != pages.size()) throw exec_err("cannot write result pages");
      std::fclose(f);
      std::printf("{\"query\": \"q1_result_pages\", \"pages\": %" PRIu64 "}\n", npages);
    }
    // outside the fused grammar (a 15-byte string literal needs the out-of-line view path): the
    // aggregate stays a DataFusion node and has no operator here
    {
      PlanRef li = plans::scan(1, plans::lineitem_q6());
      PlanRef agg = std::make_shared<AggregateExec>(
          AggregateMode::Single, std::vector<std::pair<ExprRef, std::string>>{}, std::vector<AggregateFunctionExpr>{count_star("n")},
          plans::filter(binary(col("l_shipdate", 3), Operator::Lt, lit("1995-01-01 00:00")), li));
      std::vector<std::string> why;
      PlanRef out = install_b200_operators(agg, &gpu, &why);
      bool not_implemented = false;
      try {
        TaskContext tc(&gpu);
        out->execute(0, tc);
      } catch (const DataFusionError& e) {
        not_implemented = e.kind() == ErrorKind::NotImplemented;
      }
      std::printf("{\"query\": \"ineligible\", \"kept\": \"%s\", \"reasons\": %zu, \"not_implemented\": %s}\n", out->name().c_str(), why.size(),
                  not_implemented ? "true" : "false");
    }
    // errors of the library surface as DataFusionError::Execution with its status, not as aborts
    {
      PlanRef missing = install_b200_operators(plans::q6(99), nullptr);  // scan 99 was never declared
      int status = 0;
      try {
        TaskContext tc(&gpu);
        missing->execute(0, tc);
      } catch (const DataFusionError& e) {
        status = e.kind() == ErrorKind::Execution ? int(e.status()) : -1;
      }
      std::printf("{\"query\": \"unknown_scan\", \"status\": %d}\n", status);
    }
  } catch (const DataFusionError& e) {
    std::fprintf(stderr, "error (%d, status %d): %s\n", int(e.kind()), int(e.status()), e.what());
    return 1;
  }
  std::printf("{\"query\": \"done\"}\n");
  return 0;
}
