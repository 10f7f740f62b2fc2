// Builds DataFusion-shaped plans from a line-oriented description on stdin, rewrites them with
// install_b200_operators (no device) and prints the lowered pgf_pipeline of each as hex, so that
// tests/test_cpp_host.py can compare randomly generated plans with the Python builder's PODs.
//
//   scan <scan_id> <ncols> <type_tag>...
//   term <col> <lt|le|gt|ge|eq|ne> <f|i|d|s> <value> [flip]     flip: written as <literal> <op'> <column>
//   key <col>
//   agg <sum|avg|count|countstar> <nfactors> { <kind 0|1|2> <col> <f|i|d> <const> }...
//   sort <is_agg> <index> <desc> <nulls_first>
//   limit <n>
//   end                                                          -> "POD <hex>" or "SKIP <reason>"
#include <array>
#include <cstdio>
#include <iostream>
#include <sstream>

#include "plans.hpp"

using namespace pgf_b200;

namespace {

ExprRef literal(const std::string& kind, const std::string& value) {
  if (kind == "f") return lit(std::stod(value));
  if (kind == "i") return lit(int64_t(std::stoll(value)));
  if (kind == "d") return lit(ScalarValue::decimal128(int64_t(std::stoll(value))));
  return lit(value.c_str());
}

Operator flipped(Operator op) {
  switch (op) {
    case Operator::Lt: return Operator::Gt;
    case Operator::LtEq: return Operator::GtEq;
    case Operator::Gt: return Operator::Lt;
    case Operator::GtEq: return Operator::LtEq;
    default: return op;
  }
}

struct Spec {
  uint64_t scan_id = 0;
  Schema schema;
  std::vector<ExprRef> terms;
  std::vector<std::pair<ExprRef, std::string>> keys;
  std::vector<AggregateFunctionExpr> aggs;
  std::vector<std::array<int, 4>> sort;
  uint64_t limit = 0;
};

ExprRef column(const Spec& s, size_t c) { return col(s.schema.field(c).name, c); }

void emit(const Spec& s) {
  PlanRef plan = std::make_shared<WorkerPgScanExec>(s.scan_id, s.schema);
  if (!s.terms.empty()) {
    ExprRef pred = s.terms[0];
    for (size_t i = 1; i < s.terms.size(); ++i) pred = and_(pred, s.terms[i]);
    plan = plans::filter(pred, plan);
  }
  plan = std::make_shared<AggregateExec>(AggregateMode::Single, s.keys, s.aggs, plan);
  if (!s.sort.empty()) {
    std::vector<PhysicalSortExpr> order;
    for (const auto& t : s.sort) {
      const size_t out_index = t[0] ? s.keys.size() + size_t(t[1]) : size_t(t[1]);
      order.push_back(PhysicalSortExpr{col(plan->schema().field(out_index).name, out_index), t[2] != 0, t[3] != 0});
    }
    plan = s.limit ? std::make_shared<SortExec>(order, plan, s.limit) : std::make_shared<SortExec>(order, plan);
  }
  std::vector<std::string> why;
  PlanRef out = install_b200_operators(plan, nullptr, &why);
  auto pipe = out->downcast<B200PipelineExec>();
  if (!pipe) {
    std::printf("SKIP %s\n", why.empty() ? "?" : why[0].c_str());
    return;
  }
  std::printf("POD ");
  const auto* bytes = reinterpret_cast<const unsigned char*>(&pipe->pod());
  for (size_t b = 0; b < sizeof(pgf_pipeline); ++b) std::printf("%02x", bytes[b]);
  std::printf("\n");
}

}  // namespace

int main() {
  Spec s;
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream in(line);
    std::string cmd;
    in >> cmd;
    try {
      if (cmd == "scan") {
        s = Spec{};
        size_t ncols = 0;
        in >> s.scan_id >> ncols;
        for (size_t c = 0; c < ncols; ++c) {
          int32_t tag = 0;
          in >> tag;
          s.schema.fields.push_back(Field{"c" + std::to_string(c), tag, false});
        }
      } else if (cmd == "term") {
        size_t c;
        std::string op, kind, value, flip;
        in >> c >> op >> kind >> value >> flip;
        const Operator o = op == "lt" ? Operator::Lt : op == "le" ? Operator::LtEq : op == "gt" ? Operator::Gt
                         : op == "ge" ? Operator::GtEq : op == "eq" ? Operator::Eq : Operator::NotEq;
        s.terms.push_back(flip == "flip" ? binary(literal(kind, value), flipped(o), column(s, c)) : binary(column(s, c), o, literal(kind, value)));
      } else if (cmd == "key") {
        size_t c;
        in >> c;
        s.keys.push_back({column(s, c), s.schema.field(c).name});
      } else if (cmd == "agg") {
        std::string func;
        size_t nf = 0;
        in >> func >> nf;
        ExprRef e;
        for (size_t f = 0; f < nf; ++f) {
          int kind;
          size_t c;
          std::string lk, lv;
          in >> kind >> c >> lk >> lv;
          ExprRef factor = kind == 0 ? column(s, c)
                         : kind == 1 ? binary(literal(lk, lv), Operator::Minus, column(s, c))
                         : (f % 2 ? binary(column(s, c), Operator::Plus, literal(lk, lv))      // both spellings of (c + x)
                                  : binary(literal(lk, lv), Operator::Plus, column(s, c)));
          e = e ? binary(e, Operator::Multiply, factor) : factor;
        }
        const std::string name = "a" + std::to_string(s.aggs.size());
        if (func == "countstar") s.aggs.push_back(count_star(name));
        else if (func == "sum") s.aggs.push_back(sum(e, name));
        else if (func == "avg") s.aggs.push_back(avg(e, name));
        else s.aggs.push_back(count(e, name));
      } else if (cmd == "sort") {
        std::array<int, 4> t{};
        in >> t[0] >> t[1] >> t[2] >> t[3];
        s.sort.push_back(t);
      } else if (cmd == "limit") {
        in >> s.limit;
      } else if (cmd == "end") {
        emit(s);
      }
    } catch (const std::exception& e) {
      std::printf("ERROR %s\n", e.what());
    }
  }
  return 0;
}
