// pgf_b200_plan.hpp -- host side above the C ABI (include/pgf_b200.h), in C++17, header only.
//
// pg_fusion's worker plans a DataFusion physical plan and rewrites it bottom-up
// (worker_runtime/src/runtime.rs:667-698: install_runtime_filters, insert_page_materializers).
// The GPU path is one more rewrite of the same shape.  The reference's host code is Rust, which
// this image cannot build; this header is the same host logic in C++ so that it is compiled,
// linked against libpgf_b200.so and tested here (tests/cpp/), and reads like the Rust shim of
// INTEGRATION.md:
//
//   * the plan-node surface of DataFusion 44 the reference's own nodes implement
//     (worker_runtime/src/runtime_filter_plan.rs:154-217, scan_exec.rs:195-268,
//     pg/scan_node/src/page_materialize.rs:43-100): name / schema / children /
//     with_new_children / execute(partition) / DisplayAs, one partition, partition != 0 rejected;
//   * descriptor nodes for the operators on the path (WorkerPgScanExec, FilterExec,
//     CoalesceBatchesExec, ProjectionExec, AggregateExec, HashJoinExec, RuntimeFilterBuildExec,
//     SortExec, GlobalLimitExec) and PhysicalExpr (Column, Literal, BinaryExpr);
//   * install_runtime_filters: the eligibility rules of runtime_filter_plan.rs:50-111;
//   * install_b200_operators: match_pipeline + lower_to_pod -> B200PipelineExec, whose execute()
//     runs build sides in dependency order and then the fused pipeline (pgf_pipeline_run).
//
// There are no CPU operators: executing a node that was not absorbed by a B200PipelineExec fails
// with ErrorKind::NotImplemented (in the reference that node would stay a DataFusion node).
// Errors are exceptions of one type, DataFusionError, with the variants the reference uses
// (Plan / Execution / External; runtime_filter_plan.rs:87,186,240); nothing here aborts.
#ifndef PGF_B200_PLAN_HPP
#define PGF_B200_PLAN_HPP

#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pgf_b200.h"

namespace pgf_b200 {

// ------------------------------------------------------------------------------------ errors
enum class ErrorKind { Plan, Execution, NotImplemented, External };

class DataFusionError : public std::runtime_error {
 public:
  DataFusionError(ErrorKind kind, const std::string& msg, pgf_status status = PGF_OK)
      : std::runtime_error(msg), kind_(kind), status_(status) {}
  ErrorKind kind() const { return kind_; }
  pgf_status status() const { return status_; }  // the C ABI status for Execution errors raised by the library

 private:
  ErrorKind kind_;
  pgf_status status_;
};
inline DataFusionError plan_err(const std::string& m) { return DataFusionError(ErrorKind::Plan, m); }
inline DataFusionError exec_err(const std::string& m, pgf_status st = PGF_OK) { return DataFusionError(ErrorKind::Execution, m, st); }

// Thrown inside the lowering only: the sub-tree is outside the fused grammar and stays as it is
// (maybe_wrap_hash_join returning Ok(None), runtime_filter_plan.rs:56-80).
struct NotEligible {
  std::string why;
};

// ------------------------------------------------------------------------------------ schema
struct Field {
  std::string name;
  int32_t type_tag = 0;  // PGF_T_*
  bool nullable = false;
};
struct Schema {
  std::vector<Field> fields;
  size_t size() const { return fields.size(); }
  const Field& field(size_t i) const {
    if (i >= fields.size()) throw plan_err("column index " + std::to_string(i) + " out of range");
    return fields[i];
  }
};

// ------------------------------------------------------------------------------ PhysicalExpr
struct ScalarValue {
  enum Kind { Null, Float64, Int64, Utf8, Decimal128, Boolean } kind = Null;
  double f64 = 0;
  int64_t lo = 0, hi = 0;  // Int64 in lo; Decimal128 unscaled value in (hi:lo)
  std::string str;
  static ScalarValue float64(double v) { ScalarValue s; s.kind = Float64; s.f64 = v; return s; }
  static ScalarValue int64(int64_t v) { ScalarValue s; s.kind = Int64; s.lo = v; s.hi = v < 0 ? -1 : 0; return s; }
  static ScalarValue utf8(std::string v) { ScalarValue s; s.kind = Utf8; s.str = std::move(v); return s; }
  static ScalarValue boolean(bool v) { ScalarValue s; s.kind = Boolean; s.lo = v ? 1 : 0; return s; }
  static ScalarValue decimal128(int64_t unscaled) { ScalarValue s; s.kind = Decimal128; s.lo = unscaled; s.hi = unscaled < 0 ? -1 : 0; return s; }
  static ScalarValue decimal128(int64_t hi, uint64_t lo) { ScalarValue s; s.kind = Decimal128; s.lo = int64_t(lo); s.hi = hi; return s; }
  std::string to_string() const {
    std::ostringstream o;
    switch (kind) {
      case Null: o << "NULL"; break;
      case Float64: o << f64; break;
      case Int64: o << lo; break;
      case Utf8: o << '"' << str << '"'; break;
      case Decimal128: o << "Decimal128(" << hi << ":" << uint64_t(lo) << ")"; break;
      case Boolean: o << (lo ? "true" : "false"); break;
    }
    return o.str();
  }
};

enum class Operator { And, Or, Lt, LtEq, Gt, GtEq, Eq, NotEq, Plus, Minus, Multiply };
inline const char* operator_str(Operator op) {
  switch (op) {
    case Operator::And: return "AND";
    case Operator::Or: return "OR";
    case Operator::Lt: return "<";
    case Operator::LtEq: return "<=";
    case Operator::Gt: return ">";
    case Operator::GtEq: return ">=";
    case Operator::Eq: return "=";
    case Operator::NotEq: return "!=";
    case Operator::Plus: return "+";
    case Operator::Minus: return "-";
    case Operator::Multiply: return "*";
  }
  return "?";
}

class PhysicalExpr {
 public:
  virtual ~PhysicalExpr() = default;
  virtual std::string to_string() const = 0;
  virtual int32_t data_type(const Schema& input) const = 0;  // PGF_T_*
  template <class T>
  const T* downcast() const { return dynamic_cast<const T*>(this); }  // as_any().downcast_ref::<T>()
};
using ExprRef = std::shared_ptr<const PhysicalExpr>;

class Column final : public PhysicalExpr {
 public:
  Column(std::string name, size_t index) : name_(std::move(name)), index_(index) {}
  const std::string& name() const { return name_; }
  size_t index() const { return index_; }
  std::string to_string() const override { return name_ + "@" + std::to_string(index_); }
  int32_t data_type(const Schema& input) const override { return input.field(index_).type_tag; }

 private:
  std::string name_;
  size_t index_;
};

class Literal final : public PhysicalExpr {
 public:
  explicit Literal(ScalarValue v) : value_(std::move(v)) {}
  const ScalarValue& value() const { return value_; }
  std::string to_string() const override { return value_.to_string(); }
  int32_t data_type(const Schema&) const override {
    switch (value_.kind) {
      case ScalarValue::Float64: return PGF_T_FLOAT64;
      case ScalarValue::Int64: return PGF_T_INT64;
      case ScalarValue::Utf8: return PGF_T_UTF8VIEW;
      case ScalarValue::Decimal128: return PGF_T_DECIMAL128;
      case ScalarValue::Boolean: return PGF_T_BOOLEAN;
      default: return 0;
    }
  }

 private:
  ScalarValue value_;
};

class BinaryExpr final : public PhysicalExpr {
 public:
  BinaryExpr(ExprRef l, Operator op, ExprRef r) : l_(std::move(l)), op_(op), r_(std::move(r)) {}
  const ExprRef& left() const { return l_; }
  const ExprRef& right() const { return r_; }
  Operator op() const { return op_; }
  std::string to_string() const override { return l_->to_string() + " " + operator_str(op_) + " " + r_->to_string(); }
  int32_t data_type(const Schema& input) const override {
    switch (op_) {
      case Operator::Plus: case Operator::Minus: case Operator::Multiply: {
        // a column decides the type (literals are coerced to it by the planner)
        const int32_t lt = l_->data_type(input), rt = r_->data_type(input);
        return l_->downcast<Literal>() ? rt : lt;
      }
      default: return PGF_T_BOOLEAN;
    }
  }

 private:
  ExprRef l_;
  Operator op_;
  ExprRef r_;
};

inline ExprRef col(const std::string& name, size_t index) { return std::make_shared<Column>(name, index); }
inline ExprRef lit(double v) { return std::make_shared<Literal>(ScalarValue::float64(v)); }
inline ExprRef lit(int64_t v) { return std::make_shared<Literal>(ScalarValue::int64(v)); }
inline ExprRef lit(const char* v) { return std::make_shared<Literal>(ScalarValue::utf8(v)); }
inline ExprRef lit(ScalarValue v) { return std::make_shared<Literal>(std::move(v)); }
inline ExprRef lit_bool(bool v) { return std::make_shared<Literal>(ScalarValue::boolean(v)); }
inline ExprRef binary(ExprRef l, Operator op, ExprRef r) { return std::make_shared<BinaryExpr>(std::move(l), op, std::move(r)); }
inline ExprRef and_(ExprRef l, ExprRef r) { return binary(std::move(l), Operator::And, std::move(r)); }

// ----------------------------------------------------------------------- device context (RAII)
struct RuntimeFilterTarget {  // runtime_filter/src/pool.rs RuntimeFilterTarget
  uint64_t session_epoch = 0, scan_id = 0;
  uint32_t output_column = 0, key_type = 0;  // 1 = Int16, 2 = Int32, 3 = Int64
};
struct RuntimeFilterBuildHandle {
  uint64_t bloom = 0;       // library handle of the filter in HBM
  uint64_t generation = 0;  // generation of the Building state this handle owns
};
// RuntimeFilterPool::allocate_build (pool.rs:378-430): nullopt = pool exhausted, a soft miss.
class RuntimeFilterPool {
 public:
  virtual ~RuntimeFilterPool() = default;
  virtual std::optional<RuntimeFilterBuildHandle> allocate_build(const RuntimeFilterTarget& target) = 0;
};

class B200Context {
 public:
  explicit B200Context(int32_t device = 0, uint32_t page_size = 65536, uint32_t flags = 0) {
    pgf_config cfg{device, page_size, 0, flags};
    const pgf_status st = pgf_ctx_create(&cfg, &ctx_);
    if (st != PGF_OK) {
      ctx_ = nullptr;
      throw exec_err("pgf_ctx_create: status " + std::to_string(st) + " (a CUDA device is required; there is no CPU fallback)", st);
    }
    page_size_ = page_size;
  }
  ~B200Context() { if (ctx_) pgf_ctx_destroy(ctx_); }
  B200Context(const B200Context&) = delete;
  B200Context& operator=(const B200Context&) = delete;
  pgf_ctx* raw() const { return ctx_; }
  uint32_t page_size() const { return page_size_; }
  // ffi::check of INTEGRATION.md: status -> DataFusionError::Execution carrying pgf_last_error
  void check(pgf_status st) const {
    if (st == PGF_OK) return;
    const char* m = pgf_last_error(ctx_);
    throw exec_err("pgf_b200 status " + std::to_string(st) + ": " + (m ? m : ""), st);
  }

 private:
  pgf_ctx* ctx_ = nullptr;
  uint32_t page_size_ = 65536;
};

// Filters in HBM: pgf_bloom_create + try_acquire_builder per allocation, at most slot_count of them
// (pg_fusion.runtime_filter_slots, pg/extension/src/guc.rs:41-46).
class DeviceRuntimeFilterPool final : public RuntimeFilterPool {
 public:
  DeviceRuntimeFilterPool(B200Context& gpu, pgf_bloom_params params, uint32_t slot_count = 64)
      : gpu_(gpu), params_(params), slots_(slot_count) {}
  ~DeviceRuntimeFilterPool() override {
    for (uint64_t b : blooms_) pgf_bloom_destroy(gpu_.raw(), b);
  }
  std::optional<RuntimeFilterBuildHandle> allocate_build(const RuntimeFilterTarget&) override {
    if (blooms_.size() >= slots_) {   // RuntimeFilterPoolExhaustedTotal (runtime_filter_plan.rs:89): a soft miss
      pgf_ctx_note_pool_exhausted(gpu_.raw());
      return std::nullopt;
    }
    RuntimeFilterBuildHandle h;
    gpu_.check(pgf_bloom_create(gpu_.raw(), &params_, &h.bloom));
    blooms_.push_back(h.bloom);
    gpu_.check(pgf_bloom_begin_build(gpu_.raw(), h.bloom, &h.generation));
    return h;
  }

 private:
  B200Context& gpu_;
  pgf_bloom_params params_;
  uint32_t slots_;
  std::vector<uint64_t> blooms_;
};

// The producer side of a scan: WorkerPgScanExec's scan thread receives transfer pages and hands
// them on (worker_runtime/src/transport_scan_source.rs:251-426, one OS thread per scan); here the
// pages go to HBM.  push_* of different scans may run concurrently; each call copies (or DMAs,
// for registered memory) before it returns, so the caller can release the page immediately --
// the deep copy PageMaterializeExec makes for retaining operators (page_materialize.rs:107-207).
class ScanIngest {
 public:
  ScanIngest(B200Context& gpu, uint64_t scan_id, const Schema& schema, uint64_t expected_pages = 0) : gpu_(gpu), scan_id_(scan_id) {
    std::vector<pgf_column_spec> specs;
    for (const auto& f : schema.fields) specs.push_back(pgf_column_spec{uint16_t(f.type_tag), uint16_t(f.nullable ? 1 : 0)});
    gpu_.check(pgf_scan_declare(gpu_.raw(), scan_id_, specs.data(), uint32_t(specs.size()), expected_pages));
  }
  void push_page(const uint8_t* page, uint32_t len) { gpu_.check(pgf_scan_push_page(gpu_.raw(), scan_id_, page, len)); }
  void push_pages(const uint8_t* pages, uint64_t npages, uint64_t stride) {
    gpu_.check(pgf_scan_push_pages(gpu_.raw(), scan_id_, pages, npages, stride));
  }
  // end of stream: waits for the copies, runs the row-level import checks on the device
  void finish() { gpu_.check(pgf_scan_finish(gpu_.raw(), scan_id_)); }
  pgf_scan_info info() const {
    pgf_scan_info i;
    gpu_.check(pgf_scan_get_info(gpu_.raw(), scan_id_, &i));
    return i;
  }
  uint64_t scan_id() const { return scan_id_; }

 private:
  B200Context& gpu_;
  uint64_t scan_id_;
};

// ------------------------------------------------------------------------------ RecordBatch
struct PipelineMetrics {  // EXPLAIN ANALYZE counters of one fused pipeline
  std::string node;
  uint64_t rows_in = 0, rows_bloom = 0, rows_filtered = 0, rows_out = 0, bloom_rows = 0;
  float kernel_ms = 0;
  uint32_t kernel_launches = 0;
  std::string variant;
};

struct RecordBatch {
  Schema schema;
  uint64_t num_rows = 0;
  std::vector<std::vector<pgf_value>> columns;  // columns[c][row]
  uint64_t join_table = 0;                      // build-side pipelines: the table handle
  std::shared_ptr<pgf_result> raw;              // kept for pgf_result_encode_pages
};

class TaskContext {
 public:
  explicit TaskContext(B200Context* gpu) : gpu_(gpu) {}
  B200Context& gpu() const {
    if (!gpu_) throw exec_err("TaskContext carries no device context (there is no CPU fallback)", PGF_ERR_NO_DEVICE);
    return *gpu_;
  }
  std::vector<PipelineMetrics> metrics;

 private:
  B200Context* gpu_;
};

// ---------------------------------------------------------------------------- ExecutionPlan
class ExecutionPlan;
using PlanRef = std::shared_ptr<const ExecutionPlan>;

class ExecutionPlan : public std::enable_shared_from_this<ExecutionPlan> {
 public:
  virtual ~ExecutionPlan() = default;
  virtual std::string name() const = 0;
  virtual const Schema& schema() const = 0;
  virtual std::vector<PlanRef> children() const = 0;
  virtual PlanRef with_new_children(std::vector<PlanRef> children) const = 0;
  virtual size_t partition_count() const { return 1; }  // Partitioning::UnknownPartitioning(1)
  virtual std::string fmt_as() const { return name(); }  // DisplayAs, one line
  virtual RecordBatch execute(size_t, TaskContext&) const {
    throw DataFusionError(ErrorKind::NotImplemented,
                          name() + ": this build has no CPU operators; only nodes absorbed by a B200PipelineExec execute");
  }
  template <class T>
  const T* downcast() const { return dynamic_cast<const T*>(this); }

 protected:
  static void expect_children(const std::string& who, const std::vector<PlanRef>& c, size_t n) {
    if (c.size() != n) throw plan_err(who + " expects " + std::to_string(n) + " child(ren), got " + std::to_string(c.size()));
  }
};

// displayable(plan).indent()
inline void display_indent(const PlanRef& plan, std::ostream& out, int depth = 0) {
  out << std::string(size_t(depth) * 2, ' ') << plan->fmt_as() << "\n";
  for (const auto& c : plan->children()) display_indent(c, out, depth + 1);
}
inline std::string display_indent(const PlanRef& plan) {
  std::ostringstream o;
  display_indent(plan, o);
  return o.str();
}

struct RuntimeFilterProbe {  // what lookup_probes(session_epoch, scan_id) yields for a scan (pool.rs:432-476)
  uint64_t bloom = 0, generation = 0;
  uint32_t output_column = 0;
};

// WorkerPgScanExec (worker_runtime/src/scan_exec.rs:139-268): the leaf.  Its pages are pushed into
// HBM through pgf_scan_push_page(s) under `scan_id`; the node itself only names them.
class WorkerPgScanExec final : public ExecutionPlan {
 public:
  WorkerPgScanExec(uint64_t scan_id, Schema schema, std::vector<RuntimeFilterProbe> probes = {})
      : scan_id_(scan_id), schema_(std::move(schema)), probes_(std::move(probes)) {}
  uint64_t scan_id() const { return scan_id_; }
  const std::vector<RuntimeFilterProbe>& runtime_filter_probes() const { return probes_; }
  PlanRef with_runtime_filter_probe(RuntimeFilterProbe p) const {
    auto probes = probes_;
    probes.push_back(p);
    return std::make_shared<WorkerPgScanExec>(scan_id_, schema_, std::move(probes));
  }
  std::string name() const override { return "WorkerPgScanExec"; }
  const Schema& schema() const override { return schema_; }
  std::vector<PlanRef> children() const override { return {}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 0);
    return shared_from_this();
  }
  std::string fmt_as() const override {
    std::ostringstream o;
    o << "WorkerPgScanExec: scan_id=" << scan_id_;
    for (const auto& p : probes_) o << ", runtime_filter(col=" << p.output_column << ")";
    return o.str();
  }

 private:
  uint64_t scan_id_;
  Schema schema_;
  std::vector<RuntimeFilterProbe> probes_;
};

class FilterExec final : public ExecutionPlan {
 public:
  FilterExec(ExprRef predicate, PlanRef input) : predicate_(std::move(predicate)), input_(std::move(input)) {}
  const ExprRef& predicate() const { return predicate_; }
  const PlanRef& input() const { return input_; }
  std::string name() const override { return "FilterExec"; }
  const Schema& schema() const override { return input_->schema(); }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<FilterExec>(predicate_, c[0]);
  }
  std::string fmt_as() const override { return "FilterExec: " + predicate_->to_string(); }

 private:
  ExprRef predicate_;
  PlanRef input_;
};

class CoalesceBatchesExec final : public ExecutionPlan {
 public:
  explicit CoalesceBatchesExec(PlanRef input, size_t target_batch_size = 8192)
      : input_(std::move(input)), target_(target_batch_size) {}
  const PlanRef& input() const { return input_; }
  std::string name() const override { return "CoalesceBatchesExec"; }
  const Schema& schema() const override { return input_->schema(); }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<CoalesceBatchesExec>(c[0], target_);
  }
  std::string fmt_as() const override { return "CoalesceBatchesExec: target_batch_size=" + std::to_string(target_); }

 private:
  PlanRef input_;
  size_t target_;
};

class ProjectionExec final : public ExecutionPlan {
 public:
  ProjectionExec(std::vector<std::pair<ExprRef, std::string>> exprs, PlanRef input)
      : exprs_(std::move(exprs)), input_(std::move(input)) {
    for (const auto& e : exprs_) {
      Field f{e.second, e.first->data_type(input_->schema()), true};
      if (auto c = e.first->downcast<Column>()) f.nullable = input_->schema().field(c->index()).nullable;
      schema_.fields.push_back(f);
    }
  }
  const std::vector<std::pair<ExprRef, std::string>>& expr() const { return exprs_; }
  const PlanRef& input() const { return input_; }
  std::string name() const override { return "ProjectionExec"; }
  const Schema& schema() const override { return schema_; }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<ProjectionExec>(exprs_, c[0]);
  }
  std::string fmt_as() const override {
    std::string s = "ProjectionExec: expr=[";
    for (size_t i = 0; i < exprs_.size(); ++i) s += (i ? ", " : "") + exprs_[i].first->to_string() + " as " + exprs_[i].second;
    return s + "]";
  }

 private:
  std::vector<std::pair<ExprRef, std::string>> exprs_;
  PlanRef input_;
  Schema schema_;
};

enum class AggregateMode { Partial, Final, FinalPartitioned, Single, SinglePartitioned };
enum class AggregateFunction { Sum, Avg, Count };
struct AggregateFunctionExpr {
  AggregateFunction fun;
  std::vector<ExprRef> args;  // COUNT(*) is count(Int64(1)) in DataFusion: a non-null literal argument (or none)
  std::string name;
};
inline AggregateFunctionExpr sum(ExprRef e, std::string name) { return {AggregateFunction::Sum, {std::move(e)}, std::move(name)}; }
inline AggregateFunctionExpr avg(ExprRef e, std::string name) { return {AggregateFunction::Avg, {std::move(e)}, std::move(name)}; }
inline AggregateFunctionExpr count(ExprRef e, std::string name) { return {AggregateFunction::Count, {std::move(e)}, std::move(name)}; }
inline AggregateFunctionExpr count_star(std::string name) { return {AggregateFunction::Count, {lit(int64_t(1))}, std::move(name)}; }

class AggregateExec final : public ExecutionPlan {
 public:
  AggregateExec(AggregateMode mode, std::vector<std::pair<ExprRef, std::string>> group_by,
                std::vector<AggregateFunctionExpr> aggr_expr, PlanRef input)
      : mode_(mode), group_by_(std::move(group_by)), aggr_(std::move(aggr_expr)), input_(std::move(input)) {
    const Schema& in = input_->schema();
    for (const auto& g : group_by_) {
      // a group field keeps the nullability of its input field (DataFusion derives it with PhysicalExpr::nullable)
      Field f{g.second, g.first->data_type(in), true};
      if (auto c = g.first->downcast<Column>()) f.nullable = in.field(c->index()).nullable;
      schema_.fields.push_back(f);
    }
    for (const auto& a : aggr_) {
      int32_t t = PGF_T_INT64;
      if (a.fun != AggregateFunction::Count) {
        const int32_t at = a.args.empty() ? 0 : a.args[0]->data_type(in);
        // SUM/AVG(Float64), AVG(int) -> Float64; SUM(int) -> Int64; SUM/AVG(Decimal128) -> Decimal128
        if (at == PGF_T_DECIMAL128) t = PGF_T_DECIMAL128;
        else if (at == PGF_T_FLOAT64 || at == PGF_T_FLOAT32 || a.fun == AggregateFunction::Avg) t = PGF_T_FLOAT64;
      }
      schema_.fields.push_back(Field{a.name, t, a.fun != AggregateFunction::Count});
    }
  }
  AggregateMode mode() const { return mode_; }
  const std::vector<std::pair<ExprRef, std::string>>& group_expr() const { return group_by_; }
  const std::vector<AggregateFunctionExpr>& aggr_expr() const { return aggr_; }
  const PlanRef& input() const { return input_; }
  std::string name() const override { return "AggregateExec"; }
  const Schema& schema() const override { return schema_; }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<AggregateExec>(mode_, group_by_, aggr_, c[0]);
  }
  std::string fmt_as() const override {
    static const char* modes[] = {"Partial", "Final", "FinalPartitioned", "Single", "SinglePartitioned"};
    std::string s = std::string("AggregateExec: mode=") + modes[int(mode_)] + ", gby=[";
    for (size_t i = 0; i < group_by_.size(); ++i) s += (i ? ", " : "") + group_by_[i].first->to_string() + " as " + group_by_[i].second;
    s += "], aggr=[";
    for (size_t i = 0; i < aggr_.size(); ++i) s += (i ? ", " : "") + aggr_[i].name;
    return s + "]";
  }

 private:
  AggregateMode mode_;
  std::vector<std::pair<ExprRef, std::string>> group_by_;
  std::vector<AggregateFunctionExpr> aggr_;
  PlanRef input_;
  Schema schema_;
};

enum class JoinType { Inner, Left, Right, Full, LeftSemi, RightSemi, LeftAnti, RightAnti };
enum class PartitionMode { Partitioned, CollectLeft, Auto };

class HashJoinExec final : public ExecutionPlan {
 public:
  using JoinOn = std::vector<std::pair<ExprRef, ExprRef>>;
  HashJoinExec(PlanRef left, PlanRef right, JoinOn on, JoinType join_type = JoinType::Inner,
               PartitionMode mode = PartitionMode::CollectLeft, bool null_equals_null = false,
               ExprRef filter = nullptr, std::optional<std::vector<size_t>> projection = std::nullopt)
      : left_(std::move(left)), right_(std::move(right)), on_(std::move(on)), join_type_(join_type), mode_(mode),
        null_equals_null_(null_equals_null), filter_(std::move(filter)), projection_(std::move(projection)) {
    Schema all;
    for (const auto& f : left_->schema().fields) all.fields.push_back(f);
    for (const auto& f : right_->schema().fields) all.fields.push_back(f);
    if (projection_) {
      for (size_t i : *projection_) schema_.fields.push_back(all.field(i));
    } else {
      schema_ = all;
    }
  }
  const PlanRef& left() const { return left_; }
  const PlanRef& right() const { return right_; }
  const JoinOn& on() const { return on_; }
  JoinType join_type() const { return join_type_; }
  PartitionMode partition_mode() const { return mode_; }
  bool null_equals_null() const { return null_equals_null_; }
  const ExprRef& filter() const { return filter_; }
  const std::optional<std::vector<size_t>>& projection() const { return projection_; }
  std::string name() const override { return "HashJoinExec"; }
  const Schema& schema() const override { return schema_; }
  std::vector<PlanRef> children() const override { return {left_, right_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 2);
    return std::make_shared<HashJoinExec>(c[0], c[1], on_, join_type_, mode_, null_equals_null_, filter_, projection_);
  }
  std::string fmt_as() const override {
    std::string s = std::string("HashJoinExec: mode=") + (mode_ == PartitionMode::CollectLeft ? "CollectLeft" : mode_ == PartitionMode::Partitioned ? "Partitioned" : "Auto") +
                    ", join_type=" + (join_type_ == JoinType::Inner ? "Inner" : "Other") + ", on=[";
    for (size_t i = 0; i < on_.size(); ++i) s += (i ? ", " : "") + ("(" + on_[i].first->to_string() + ", " + on_[i].second->to_string() + ")");
    return s + "]";
  }

 private:
  PlanRef left_, right_;
  JoinOn on_;
  JoinType join_type_;
  PartitionMode mode_;
  bool null_equals_null_;
  ExprRef filter_;
  std::optional<std::vector<size_t>> projection_;
  Schema schema_;
};

// RuntimeFilterBuildExec (runtime_filter_plan.rs:119-217): pass-through of the build side that
// inserts every non-null key and publishes Ready at end of stream.  Here it is absorbed by the
// build-side pipeline (pgf_pipeline.build_bloom).
class RuntimeFilterBuildExec final : public ExecutionPlan {
 public:
  RuntimeFilterBuildExec(PlanRef input, size_t key_index, uint32_t key_type, RuntimeFilterBuildHandle handle)
      : input_(std::move(input)), key_index_(key_index), key_type_(key_type), handle_(handle) {}
  const PlanRef& input() const { return input_; }
  size_t key_index() const { return key_index_; }
  uint32_t key_type() const { return key_type_; }
  const RuntimeFilterBuildHandle& handle() const { return handle_; }
  std::string name() const override { return "RuntimeFilterBuildExec"; }
  const Schema& schema() const override { return input_->schema(); }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<RuntimeFilterBuildExec>(c[0], key_index_, key_type_, handle_);
  }
  std::string fmt_as() const override { return "RuntimeFilterBuildExec: key_index=" + std::to_string(key_index_); }

 private:
  PlanRef input_;
  size_t key_index_;
  uint32_t key_type_;
  RuntimeFilterBuildHandle handle_;
};

struct PhysicalSortExpr {
  ExprRef expr;
  bool descending = false;
  bool nulls_first = false;  // DataFusion defaults: ASC NULLS LAST, DESC NULLS FIRST
};
inline PhysicalSortExpr sort_asc(ExprRef e) { return {std::move(e), false, false}; }
inline PhysicalSortExpr sort_desc(ExprRef e) { return {std::move(e), true, true}; }

class SortExec final : public ExecutionPlan {
 public:
  SortExec(std::vector<PhysicalSortExpr> expr, PlanRef input, std::optional<uint64_t> fetch = std::nullopt)
      : expr_(std::move(expr)), input_(std::move(input)), fetch_(fetch) {}
  const std::vector<PhysicalSortExpr>& expr() const { return expr_; }
  const PlanRef& input() const { return input_; }
  const std::optional<uint64_t>& fetch() const { return fetch_; }
  std::string name() const override { return "SortExec"; }
  const Schema& schema() const override { return input_->schema(); }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<SortExec>(expr_, c[0], fetch_);
  }
  std::string fmt_as() const override {
    std::string s = fetch_ ? "SortExec: TopK(fetch=" + std::to_string(*fetch_) + "), expr=[" : "SortExec: expr=[";
    for (size_t i = 0; i < expr_.size(); ++i) s += (i ? ", " : "") + expr_[i].expr->to_string() + (expr_[i].descending ? " DESC" : " ASC");
    return s + "]";
  }

 private:
  std::vector<PhysicalSortExpr> expr_;
  PlanRef input_;
  std::optional<uint64_t> fetch_;
};

class GlobalLimitExec final : public ExecutionPlan {
 public:
  GlobalLimitExec(PlanRef input, uint64_t skip, std::optional<uint64_t> fetch) : input_(std::move(input)), skip_(skip), fetch_(fetch) {}
  const PlanRef& input() const { return input_; }
  uint64_t skip() const { return skip_; }
  const std::optional<uint64_t>& fetch() const { return fetch_; }
  std::string name() const override { return "GlobalLimitExec"; }
  const Schema& schema() const override { return input_->schema(); }
  std::vector<PlanRef> children() const override { return {input_}; }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, 1);
    return std::make_shared<GlobalLimitExec>(c[0], skip_, fetch_);
  }
  std::string fmt_as() const override {
    return "GlobalLimitExec: skip=" + std::to_string(skip_) + ", fetch=" + (fetch_ ? std::to_string(*fetch_) : std::string("None"));
  }

 private:
  PlanRef input_;
  uint64_t skip_;
  std::optional<uint64_t> fetch_;
};

// ------------------------------------------------------------------ install_runtime_filters
inline std::optional<uint32_t> key_type_for(int32_t type_tag) {  // runtime_filter_plan.rs:113-120
  switch (type_tag) {
    case PGF_T_INT16: return 1u;
    case PGF_T_INT32: return 2u;
    case PGF_T_INT64: return 3u;
    default: return std::nullopt;
  }
}

namespace detail {
// The probe-side scan under schema-preserving nodes.  The reference requires the scan to be the
// join's direct right child (runtime_filter_plan.rs:74-76) because its predicates are pushed into
// the PostgreSQL scan; here predicates run on the GPU, so FilterExec / CoalesceBatchesExec above
// the scan (same column indices) are looked through.
inline const WorkerPgScanExec* probe_side_scan(const PlanRef& p) {
  if (auto s = p->downcast<WorkerPgScanExec>()) return s;
  if (auto f = p->downcast<FilterExec>()) return probe_side_scan(f->input());
  if (auto c = p->downcast<CoalesceBatchesExec>()) return probe_side_scan(c->input());
  return nullptr;
}
inline PlanRef replace_scan(const PlanRef& p, const WorkerPgScanExec* scan, const PlanRef& with) {
  if (p.get() == scan) return with;
  std::vector<PlanRef> kids;
  for (const auto& c : p->children()) kids.push_back(replace_scan(c, scan, with));
  return p->with_new_children(std::move(kids));
}
}  // namespace detail

// maybe_wrap_hash_join (runtime_filter_plan.rs:50-111): Inner, !null_equals_null, one-partition
// build side, exactly one (Column, Column) key pair, probe side a scan, key Int16/32/64.
inline std::optional<PlanRef> maybe_wrap_hash_join(const HashJoinExec& join, uint64_t session_epoch, RuntimeFilterPool& pool) {
  if (join.join_type() != JoinType::Inner || join.null_equals_null()) return std::nullopt;
  if (join.left()->partition_count() != 1) return std::nullopt;
  if (join.on().size() != 1) return std::nullopt;
  const Column* left_col = join.on()[0].first->downcast<Column>();
  const Column* right_col = join.on()[0].second->downcast<Column>();
  if (!left_col || !right_col) return std::nullopt;
  const WorkerPgScanExec* right_scan = detail::probe_side_scan(join.right());
  if (!right_scan) return std::nullopt;
  const auto key_type = key_type_for(join.right()->schema().field(right_col->index()).type_tag);
  if (!key_type) return std::nullopt;
  RuntimeFilterTarget target{session_epoch, right_scan->scan_id(), uint32_t(right_col->index()), *key_type};
  const auto handle = pool.allocate_build(target);
  if (!handle) return std::nullopt;  // RuntimeFilterPoolExhaustedTotal: a soft miss
  PlanRef left = std::make_shared<RuntimeFilterBuildExec>(join.left(), left_col->index(), *key_type, *handle);
  PlanRef right = detail::replace_scan(join.right(), right_scan,
                                       right_scan->with_runtime_filter_probe({handle->bloom, handle->generation, target.output_column}));
  return PlanRef(std::make_shared<HashJoinExec>(left, right, join.on(), join.join_type(), join.partition_mode(),
                                                join.null_equals_null(), join.filter(), join.projection()));
}

inline PlanRef install_runtime_filters(const PlanRef& plan, uint64_t session_epoch, RuntimeFilterPool& pool) {
  std::vector<PlanRef> kids;
  for (const auto& c : plan->children()) kids.push_back(install_runtime_filters(c, session_epoch, pool));
  PlanRef p = kids.empty() ? plan : plan->with_new_children(std::move(kids));
  if (auto join = p->downcast<HashJoinExec>()) {
    if (auto wrapped = maybe_wrap_hash_join(*join, session_epoch, pool)) return *wrapped;
  }
  return p;
}

// ------------------------------------------------------------------------ B200PipelineExec
class B200PipelineExec final : public ExecutionPlan {
 public:
  struct Output {  // one output column: group key `index` or aggregate `index` of the pod
    bool is_agg = false;
    uint32_t index = 0;
  };
  B200PipelineExec(pgf_pipeline pod, std::vector<PlanRef> builds, PlanRef scan, Schema schema, std::vector<Output> outputs,
                   int64_t groups_bounded_by_join = -1)
      : pod_(pod), builds_(std::move(builds)), scan_(std::move(scan)), schema_(std::move(schema)), outputs_(std::move(outputs)),
        groups_join_(groups_bounded_by_join) {}
  const pgf_pipeline& pod() const { return pod_; }
  const std::vector<PlanRef>& builds() const { return builds_; }
  const std::vector<Output>& outputs() const { return outputs_; }
  int64_t groups_bounded_by_join() const { return groups_join_; }
  const PlanRef& scan() const { return scan_; }

  std::string name() const override { return "B200PipelineExec"; }
  const Schema& schema() const override { return schema_; }
  // build sides first (in pgf_pipeline.joins[] order), then the scan leaf
  std::vector<PlanRef> children() const override {
    std::vector<PlanRef> c = builds_;
    c.push_back(scan_);
    return c;
  }
  PlanRef with_new_children(std::vector<PlanRef> c) const override {
    expect_children(name(), c, builds_.size() + 1);
    PlanRef scan = c.back();
    c.pop_back();
    return std::make_shared<B200PipelineExec>(pod_, std::move(c), std::move(scan), schema_, outputs_, groups_join_);
  }
  std::string fmt_as() const override {
    std::ostringstream o;
    o << "B200PipelineExec: scan_id=" << pod_.scan_id << ", bloom_probes=" << pod_.nbloom << ", terms=" << pod_.nterms
      << ", joins=" << pod_.njoins << ", sink=" << (pod_.sink == PGF_SINK_AGGREGATE ? "aggregate" : pod_.sink == PGF_SINK_JOIN_BUILD ? "join_build" : "count");
    if (pod_.sink == PGF_SINK_AGGREGATE) o << "(keys=" << pod_.nkeys << ", aggs=" << pod_.naggs << ")";
    if (pod_.sink == PGF_SINK_JOIN_BUILD) o << "(payload=" << pod_.npayload << (pod_.build_bloom ? ", runtime_filter" : "") << ")";
    if (pod_.nsort) o << ", sort=" << pod_.nsort;
    if (pod_.limit) o << ", fetch=" << pod_.limit;
    return o.str();
  }

  RecordBatch execute(size_t partition, TaskContext& tc) const override {
    if (partition != 0)  // scan_exec.rs:248-252
      throw plan_err("B200PipelineExec only exposes partition 0, got " + std::to_string(partition));
    B200Context& gpu = tc.gpu();
    pgf_pipeline pod = pod_;
    struct Tables {  // join tables of this pipeline live until it has run
      B200Context& gpu;
      std::vector<uint64_t> h;
      ~Tables() { for (uint64_t t : h) pgf_join_table_destroy(gpu.raw(), t); }
    } tables{gpu, {}};
    // HashJoinExec(CollectLeft) drains its left child first: run the build sides in order
    for (size_t j = 0; j < builds_.size(); ++j) {
      RecordBatch b = builds_[j]->execute(0, tc);
      tables.h.push_back(b.join_table);
      pod.joins[j].join_table = b.join_table;
    }
    if (pod.sink == PGF_SINK_AGGREGATE && pod.expected_groups == 0 && groups_join_ >= 0) {
      // a group key is the probe key of join `groups_join_`: at most one group per build row
      pgf_join_info info;
      gpu.check(pgf_join_table_get_info(gpu.raw(), tables.h[size_t(groups_join_)], &info));
      pod.expected_groups = info.rows < 1024 ? 1024 : info.rows;
    }
    pgf_result* res = nullptr;
    const pgf_status st = pgf_pipeline_run(gpu.raw(), &pod, &res);
    if (st != PGF_OK) {
      // RuntimeFilterBuildStream disables its filter on error (runtime_filter_plan.rs:309-337)
      if (pod.build_bloom) pgf_bloom_disable_build(gpu.raw(), pod.build_bloom);
      gpu.check(st);
    }
    std::shared_ptr<pgf_result> raw(res, [](pgf_result* r) { pgf_result_free(r); });
    if (pod.build_bloom) gpu.check(pgf_bloom_publish_ready(gpu.raw(), pod.build_bloom));  // Ready at end of stream
    PipelineMetrics m;
    m.node = fmt_as();
    m.rows_in = res->rows_in; m.rows_bloom = res->rows_bloom; m.rows_filtered = res->rows_filtered; m.rows_out = res->rows_out;
    m.bloom_rows = res->bloom_rows; m.kernel_ms = res->kernel_ms; m.kernel_launches = res->kernel_launches;
    m.variant = std::string(res->variant, strnlen(res->variant, sizeof res->variant));
    tc.metrics.push_back(m);
    RecordBatch out;
    out.schema = schema_;
    out.raw = raw;
    if (pod.sink == PGF_SINK_JOIN_BUILD) {
      out.join_table = res->join_table;
      out.num_rows = res->rows_out;
      return out;
    }
    out.num_rows = res->ngroups;
    out.columns.resize(outputs_.size());
    for (size_t c = 0; c < outputs_.size(); ++c) {
      out.columns[c].resize(res->ngroups);
      const Output& o = outputs_[c];
      for (uint64_t g = 0; g < res->ngroups; ++g)
        out.columns[c][g] = o.is_agg ? res->aggs[g * res->naggs + o.index] : res->keys[g * res->nkeys + o.index];
    }
    return out;
  }

 private:
  pgf_pipeline pod_;
  std::vector<PlanRef> builds_;
  PlanRef scan_;
  Schema schema_;
  std::vector<Output> outputs_;
  int64_t groups_join_;
};

// ResultPageProducer::next_step for a GPU root (worker_runtime/src/result_pages.rs:118-196): the
// result rows as transfer pages in the reference's format.  Columns are in the pod's order (group
// keys, then aggregates).
inline std::vector<uint8_t> encode_result_pages(const RecordBatch& batch, uint32_t page_size, uint64_t* npages_out = nullptr) {
  if (!batch.raw) throw exec_err("encode_result_pages: the batch does not come from a B200PipelineExec");
  pgf_column_spec schema[PGF_MAX_KEYS + PGF_MAX_AGGS];
  uint32_t ncols = 0, cap = 0;
  pgf_status st = pgf_result_schema(batch.raw.get(), schema, &ncols);
  if (st == PGF_OK) st = pgf_layout_fixed_row_cap(schema, ncols, page_size - PGF_PAGE_HEADER_LEN, &cap);
  if (st != PGF_OK || cap == 0) throw exec_err("encode_result_pages: result rows do not fit a page", st);
  const uint64_t npages = (batch.raw->ngroups + cap - 1) / cap;
  std::vector<uint8_t> pages(size_t(npages) * page_size);
  uint64_t got = 0, rows = 0;
  st = pgf_result_encode_pages(batch.raw.get(), page_size, 0, pages.data(), npages, &got, &rows);
  if (st != PGF_OK || got != npages || rows != batch.raw->ngroups) throw exec_err("pgf_result_encode_pages failed", st);
  if (npages_out) *npages_out = npages;
  return pages;
}

// ResultPageProducer::next_step as a state machine (result_pages.rs:118-148): one call produces at
// most one outbound page; after the last row one CloseFrame step; std::nullopt only after the close
// step has been handed out.  Empty results go straight to the close step (empty batches are
// skipped, result_pages.rs:137).
struct ResultPageStep {
  enum Kind { OutboundPage, CloseFrame } kind = CloseFrame;
  std::vector<uint8_t> page;  // OutboundPage: one transfer page (20-byte header + arrow_layout block)
  uint64_t rows = 0;          // rows carried by this page
};

class ResultPageProducer {
 public:
  ResultPageProducer(RecordBatch batch, uint32_t page_size) : batch_(std::move(batch)), page_size_(page_size) {
    if (!batch_.raw) throw exec_err("ResultPageProducer: the batch does not come from a B200PipelineExec");
    pgf_column_spec schema[PGF_MAX_KEYS + PGF_MAX_AGGS];
    uint32_t ncols = 0;
    pgf_status st = pgf_result_schema(batch_.raw.get(), schema, &ncols);
    if (st != PGF_OK) throw exec_err("ResultPageProducer: the result has no transport schema", st);
    specs_.assign(schema, schema + ncols);
    st = pgf_layout_fixed_row_cap(schema, ncols, page_size_ - PGF_PAGE_HEADER_LEN, &rows_per_page_);
    if (st != PGF_OK || rows_per_page_ == 0) throw exec_err("ResultPageProducer: one result row does not fit a page", st);
  }
  // normalize_result_transport_schema applied to the aggregate's output (result_pages.rs:201-249)
  const std::vector<pgf_column_spec>& transport_schema() const { return specs_; }
  uint32_t rows_per_page() const { return rows_per_page_; }

  std::optional<ResultPageStep> next_step() {
    if (pending_row_ < batch_.raw->ngroups) {
      ResultPageStep step;
      step.kind = ResultPageStep::OutboundPage;
      step.page.resize(page_size_);
      uint64_t npages = 0;
      const pgf_status st = pgf_result_encode_pages(batch_.raw.get(), page_size_, pending_row_, step.page.data(), 1, &npages, &step.rows);
      if (st != PGF_OK || npages != 1 || step.rows == 0) throw exec_err("pgf_result_encode_pages failed", st);
      pending_row_ += step.rows;
      return step;
    }
    if (close_emitted_) return std::nullopt;
    close_emitted_ = true;
    return ResultPageStep{};
  }

 private:
  RecordBatch batch_;
  uint32_t page_size_;
  std::vector<pgf_column_spec> specs_;
  uint32_t rows_per_page_ = 0;
  uint64_t pending_row_ = 0;
  bool close_emitted_ = false;
};

// ------------------------------------------------------------------ lowering (lower_to_pod)
namespace detail {

struct Bound {  // a column reference resolved down to the pipeline's sources
  pgf_colref ref{0, 0};
  int32_t type_tag = 0;
};

// An expression whose Column leaves have been resolved: BoundColumn stands for a pgf_colref.
class BoundColumn final : public PhysicalExpr {
 public:
  explicit BoundColumn(Bound b) : b_(b) {}
  const Bound& bound() const { return b_; }
  std::string to_string() const override { return "#" + std::to_string(b_.ref.source) + "." + std::to_string(b_.ref.col); }
  int32_t data_type(const Schema&) const override { return b_.type_tag; }

 private:
  Bound b_;
};

// pgf_literal for a value compared with / combined with a column of type `col_tag`
// (0 = unknown: payload columns of a join build side).  Integer literals against a Decimal128
// column are unscaled Decimal128 values.
inline pgf_literal make_literal(const ScalarValue& v, int32_t col_tag) {
  pgf_literal l;
  std::memset(&l, 0, sizeof l);
  switch (v.kind) {
    case ScalarValue::Float64:
      l.type_tag = PGF_T_FLOAT64;
      l.f64 = v.f64;
      break;
    case ScalarValue::Int64:
      l.type_tag = col_tag == PGF_T_DECIMAL128 ? PGF_T_DECIMAL128 : PGF_T_INT64;
      l.i64 = v.lo;
      l.hi = v.lo < 0 ? -1 : 0;
      break;
    case ScalarValue::Decimal128:
      l.type_tag = PGF_T_DECIMAL128;
      l.i64 = v.lo;
      l.hi = v.hi;
      break;
    case ScalarValue::Utf8:
      if (v.str.size() > 12) throw NotEligible{"string literal longer than 12 bytes (needs the out-of-line view path)"};
      l.type_tag = PGF_T_UTF8VIEW;
      l.slen = int32_t(v.str.size());
      std::memcpy(l.str, v.str.data(), v.str.size());
      break;
    case ScalarValue::Boolean:
      l.type_tag = PGF_T_BOOLEAN;
      l.i64 = v.lo;
      break;
    case ScalarValue::Null:
      throw NotEligible{"NULL literal"};
  }
  return l;
}

constexpr uint32_t kMaxFusedJoinProbes = 2;   // = PGF_MAX_JOINS: the second probe's key may be a payload column of the first

inline uint32_t value_words(int32_t type_tag) {  // 32-bit words of a value in a join-table slot
  switch (type_tag) {
    case PGF_T_INT16: case PGF_T_INT32: case PGF_T_FLOAT32: return 1;
    case PGF_T_INT64: case PGF_T_FLOAT64: return 2;
    case PGF_T_UTF8VIEW: case PGF_T_BINARYVIEW: case PGF_T_DECIMAL128: return 4;
    default: return 0;
  }
}

struct BuildSpec {
  const HashJoinExec* join = nullptr;
  std::vector<size_t> payload;  // left output columns carried in the table, in slot order
};

class Lowering {
 public:
  pgf_pipeline pod;
  std::vector<BuildSpec> builds;  // joins[] order
  PlanRef scan;

  Lowering() { std::memset(&pod, 0, sizeof pod); }

  // Walk the probe side down to the scan: predicates, joins (inner first), Bloom probes.
  void walk(const PlanRef& node) {
    if (auto s = node->downcast<WorkerPgScanExec>()) {
      if (scan) throw NotEligible{"more than one scan on the probe side"};
      scan = node;
      pod.scan_id = s->scan_id();
      for (const auto& p : s->runtime_filter_probes()) {
        if (pod.nbloom >= PGF_MAX_BLOOM_PROBES) throw NotEligible{"too many runtime filters on one scan"};
        pgf_bloom_probe& b = pod.bloom[pod.nbloom++];
        b.bloom = p.bloom;
        b.expected_generation = p.generation;
        b.key = pgf_colref{0, int32_t(p.output_column)};
      }
      return;
    }
    if (auto f = node->downcast<FilterExec>()) {
      walk(f->input());
      add_conjuncts(rebase(f->input(), f->predicate()));
      return;
    }
    if (auto c = node->downcast<CoalesceBatchesExec>()) return walk(c->input());
    if (auto p = node->downcast<ProjectionExec>()) return walk(p->input());
    if (auto j = node->downcast<HashJoinExec>()) {
      if (j->join_type() != JoinType::Inner || j->null_equals_null() || j->filter()) throw NotEligible{"join is not a plain inner equi-join"};
      if (j->partition_mode() != PartitionMode::CollectLeft) throw NotEligible{"join is not CollectLeft"};
      if (j->on().size() != 1 || !j->on()[0].first->downcast<Column>() || !j->on()[0].second->downcast<Column>())
        throw NotEligible{"join needs exactly one (column, column) key pair"};
      walk(j->right());
      // the fused kernel probes up to two join tables per pipeline (the second one in stage C, for the rows the first
      // join matched); a third probe on the same stream keeps its DataFusion node
      if (pod.njoins >= kMaxFusedJoinProbes) throw NotEligible{"more than two join probes on one scan stream"};
      const uint32_t slot = pod.njoins++;
      builds.push_back(BuildSpec{j, {}});
      const Bound key = bind(j->right(), j->on()[0].second->downcast<Column>()->index());
      if (!key_type_for(key.type_tag)) throw NotEligible{"join key is not Int16/Int32/Int64"};
      pod.joins[slot].probe_key = key.ref;
      return;
    }
    throw NotEligible{node->name() + " cannot be fused into a pipeline"};
  }

  // Output column `index` of `node` in terms of the pipeline's sources.
  Bound bind(const PlanRef& node, size_t index) {
    if (node->downcast<WorkerPgScanExec>()) return Bound{pgf_colref{0, int32_t(index)}, node->schema().field(index).type_tag};
    if (auto f = node->downcast<FilterExec>()) return bind(f->input(), index);
    if (auto c = node->downcast<CoalesceBatchesExec>()) return bind(c->input(), index);
    if (auto r = node->downcast<RuntimeFilterBuildExec>()) return bind(r->input(), index);
    if (auto p = node->downcast<ProjectionExec>()) {
      if (index >= p->expr().size()) throw plan_err("projection index out of range");
      auto c = p->expr()[index].first->downcast<Column>();
      if (!c) throw NotEligible{"computed projection column used as a key"};
      return bind(p->input(), c->index());
    }
    if (auto j = node->downcast<HashJoinExec>()) {
      size_t i = index;
      if (j->projection()) {
        if (i >= j->projection()->size()) throw plan_err("join projection index out of range");
        i = (*j->projection())[i];
      }
      const size_t nleft = j->left()->schema().size();
      if (i >= nleft) return bind(j->right(), i - nleft);
      for (size_t b = 0; b < builds.size(); ++b) {
        if (builds[b].join != j) continue;
        auto& pl = builds[b].payload;
        size_t slot = 0;
        while (slot < pl.size() && pl[slot] != i) ++slot;
        if (slot == pl.size()) {
          if (pl.size() >= PGF_MAX_PAYLOAD) throw NotEligible{"too many build-side columns used above the join"};
          pl.push_back(i);
        }
        return Bound{pgf_colref{int32_t(b) + 1, int32_t(slot)}, j->left()->schema().field(i).type_tag};
      }
      throw plan_err("join was not registered before binding");
    }
    throw NotEligible{node->name() + " cannot be fused into a pipeline"};
  }

  // Replace Column leaves (indices into `node`'s output) by BoundColumn; projections are inlined.
  ExprRef rebase(const PlanRef& node, const ExprRef& e) {
    if (auto c = e->downcast<Column>()) {
      if (auto p = node->downcast<ProjectionExec>()) {
        if (c->index() >= p->expr().size()) throw plan_err("projection index out of range");
        return rebase(p->input(), p->expr()[c->index()].first);
      }
      if (auto f = node->downcast<FilterExec>()) return rebase(f->input(), e);
      if (auto cb = node->downcast<CoalesceBatchesExec>()) return rebase(cb->input(), e);
      return std::make_shared<BoundColumn>(bind(node, c->index()));
    }
    if (auto b = e->downcast<BinaryExpr>()) return binary(rebase(node, b->left()), b->op(), rebase(node, b->right()));
    return e;  // literals
  }

  void add_conjuncts(const ExprRef& e) {
    if (auto flag = e->downcast<BoundColumn>()) {   // WHERE flag: a Boolean column is a predicate by itself (flag = true)
      if (flag->bound().type_tag != PGF_T_BOOLEAN || flag->bound().ref.source != 0) throw NotEligible{"predicate is not a conjunction of comparisons"};
      if (pod.nterms >= PGF_MAX_TERMS) throw NotEligible{"too many predicate terms"};
      pgf_pred_term& t = pod.terms[pod.nterms++];
      t.col = flag->bound().ref;
      t.cmp = PGF_CMP_EQ;
      t.lit = make_literal(ScalarValue::boolean(true), PGF_T_BOOLEAN);
      return;
    }
    auto b = e->downcast<BinaryExpr>();
    if (!b) throw NotEligible{"predicate is not a conjunction of comparisons"};
    if (b->op() == Operator::And) {
      add_conjuncts(b->left());
      add_conjuncts(b->right());
      return;
    }
    int32_t cmp;
    switch (b->op()) {
      case Operator::Lt: cmp = PGF_CMP_LT; break;
      case Operator::LtEq: cmp = PGF_CMP_LE; break;
      case Operator::Gt: cmp = PGF_CMP_GT; break;
      case Operator::GtEq: cmp = PGF_CMP_GE; break;
      case Operator::Eq: cmp = PGF_CMP_EQ; break;
      case Operator::NotEq: cmp = PGF_CMP_NE; break;
      default: throw NotEligible{std::string("operator ") + operator_str(b->op()) + " in a predicate"};
    }
    const BoundColumn* column = b->left()->downcast<BoundColumn>();
    const Literal* literal = b->right()->downcast<Literal>();
    if (!column || !literal) {  // literal <cmp> column: flip the comparison
      column = b->right()->downcast<BoundColumn>();
      literal = b->left()->downcast<Literal>();
      static const int32_t flipped[] = {PGF_CMP_GT, PGF_CMP_GE, PGF_CMP_LT, PGF_CMP_LE, PGF_CMP_EQ, PGF_CMP_NE};
      cmp = flipped[cmp];
    }
    if (!column || !literal) throw NotEligible{"comparison is not <column> <cmp> <literal>"};
    if (column->bound().ref.source != 0) throw NotEligible{"predicate on a build-side column (predicates are evaluated on scan columns)"};
    if (pod.nterms >= PGF_MAX_TERMS) throw NotEligible{"too many predicate terms"};
    pgf_pred_term& t = pod.terms[pod.nterms++];
    t.col = column->bound().ref;
    t.cmp = cmp;
    t.lit = make_literal(literal->value(), column->bound().ref.source == 0 ? column->bound().type_tag : 0);
  }

  // A product of up to three factors x, (c - x), (c + x)  (pgf_value_expr).
  void factors(const ExprRef& e, pgf_value_expr& out) {
    auto push = [&](int32_t kind, const Bound& col, const ScalarValue& c) {
      if (out.nfactors >= 3) throw NotEligible{"more than three factors in an aggregate argument"};
      pgf_factor& f = out.factors[out.nfactors++];
      f.kind = kind;
      f.col = col.ref;
      f.c = make_literal(c, col.ref.source == 0 ? col.type_tag : 0);
    };
    if (auto c = e->downcast<BoundColumn>()) return push(PGF_FACTOR_COL, c->bound(), ScalarValue::int64(0));
    auto b = e->downcast<BinaryExpr>();
    if (!b) throw NotEligible{"aggregate argument is not a product of columns"};
    if (b->op() == Operator::Multiply) {
      factors(b->left(), out);
      factors(b->right(), out);
      return;
    }
    const Literal* ll = b->left()->downcast<Literal>();
    const Literal* rl = b->right()->downcast<Literal>();
    const BoundColumn* lc = b->left()->downcast<BoundColumn>();
    const BoundColumn* rc = b->right()->downcast<BoundColumn>();
    if (b->op() == Operator::Minus && ll && rc) return push(PGF_FACTOR_CONST_MINUS_COL, rc->bound(), ll->value());
    if (b->op() == Operator::Plus && ll && rc) return push(PGF_FACTOR_CONST_PLUS_COL, rc->bound(), ll->value());
    if (b->op() == Operator::Plus && lc && rl) return push(PGF_FACTOR_CONST_PLUS_COL, lc->bound(), rl->value());
    throw NotEligible{"aggregate argument outside the x, (c - x), (c + x) product grammar"};
  }

  static void collect_types(const ExprRef& e, std::vector<int32_t>& out) {
    if (auto c = e->downcast<BoundColumn>()) out.push_back(c->bound().type_tag);
    if (auto b = e->downcast<BinaryExpr>()) {
      collect_types(b->left(), out);
      collect_types(b->right(), out);
    }
  }

  int32_t value_expr(const ExprRef& e) {
    pgf_value_expr v;
    std::memset(&v, 0, sizeof v);
    std::vector<int32_t> types;
    collect_types(e, types);
    factors(e, v);
    // one arithmetic class per expression (Float64, Decimal128 or integers); narrow integers wrap at
    // their own width in arrow, so only a plain column or all-Int64 arithmetic is computed exactly
    int cls = -1;
    bool all_i64 = true;
    for (int32_t t : types) {
      const int c = t == PGF_T_FLOAT64 ? 0 : t == PGF_T_DECIMAL128 ? 1 : key_type_for(t) ? 2 : -1;
      if (c < 0) throw NotEligible{"aggregate argument of a type the fused kernels do not compute"};
      if (cls >= 0 && cls != c) throw NotEligible{"mixed-type arithmetic in one aggregate argument"};
      cls = c;
      all_i64 = all_i64 && t == PGF_T_INT64;
    }
    if (cls == 2 && !all_i64 && !(v.nfactors == 1 && v.factors[0].kind == PGF_FACTOR_COL))
      throw NotEligible{"arithmetic on Int16/Int32 columns"};
    for (uint32_t i = 0; i < pod.nexprs; ++i)  // SUM(x) and AVG(x) share one accumulator
      if (std::memcmp(&pod.exprs[i], &v, sizeof v) == 0) return int32_t(i);
    if (pod.nexprs >= PGF_MAX_EXPRS) throw NotEligible{"too many distinct aggregate arguments"};
    pod.exprs[pod.nexprs] = v;
    return int32_t(pod.nexprs++);
  }
};

inline PlanRef lower_build(const HashJoinExec& join, const std::vector<size_t>& payload);

// Build-side pipelines of a lowered probe side, in joins[] order.
inline std::vector<PlanRef> lower_builds(const Lowering& l) {
  std::vector<PlanRef> out;
  for (const auto& b : l.builds) out.push_back(lower_build(*b.join, b.payload));
  return out;
}

// HashJoinExec.left [RuntimeFilterBuildExec] <- stream  ==>  pipeline with PGF_SINK_JOIN_BUILD
inline PlanRef lower_build(const HashJoinExec& join, const std::vector<size_t>& payload) {
  PlanRef left = join.left();
  const size_t key_index = join.on()[0].first->downcast<Column>()->index();
  Lowering l;
  if (auto rf = left->downcast<RuntimeFilterBuildExec>()) {
    if (rf->key_index() != key_index) throw NotEligible{"runtime filter key differs from the join key"};
    l.pod.build_bloom = rf->handle().bloom;
    left = rf->input();
  }
  l.walk(left);
  l.pod.sink = PGF_SINK_JOIN_BUILD;
  const Bound key = l.bind(left, key_index);
  if (key.ref.source != 0 || !key_type_for(key.type_tag)) throw NotEligible{"build key must be an Int16/Int32/Int64 column of the build-side scan"};
  l.pod.build_key = key.ref;
  Schema schema;
  schema.fields.push_back(left->schema().field(key_index));
  uint32_t payload_words = 0;  // 32-bit words carried next to the key: at most 5 (a 32-byte slot)
  for (size_t c : payload) {
    const Bound b = l.bind(left, c);
    if (b.ref.source != 0) throw NotEligible{"payload column does not come from the build-side scan"};
    payload_words += value_words(b.type_tag);
    if (value_words(b.type_tag) == 0 || payload_words > 5) throw NotEligible{"build-side columns used above the join are wider than 20 bytes"};
    l.pod.payload[l.pod.npayload++] = b.ref;
    schema.fields.push_back(left->schema().field(c));
  }
  return std::make_shared<B200PipelineExec>(l.pod, lower_builds(l), l.scan, schema, std::vector<B200PipelineExec::Output>{});
}

// AggregateExec(Single) [or Final over Partial] <- stream  ==>  pipeline with PGF_SINK_AGGREGATE
inline PlanRef lower_aggregate(const AggregateExec& agg) {
  PlanRef input = agg.input();
  if (agg.mode() == AggregateMode::Final || agg.mode() == AggregateMode::FinalPartitioned) {
    // Partial -> Final over one partition is the Single computation
    PlanRef below = input;
    if (auto cb = below->downcast<CoalesceBatchesExec>()) below = cb->input();
    auto partial = below->downcast<AggregateExec>();
    if (!partial || partial->mode() != AggregateMode::Partial || partial->aggr_expr().size() != agg.aggr_expr().size() ||
        partial->group_expr().size() != agg.group_expr().size())
      throw NotEligible{"Final aggregate without its Partial below"};
    return lower_aggregate(AggregateExec(AggregateMode::Single, partial->group_expr(), partial->aggr_expr(), partial->input()));
  }
  if (agg.mode() != AggregateMode::Single && agg.mode() != AggregateMode::SinglePartitioned) throw NotEligible{"aggregate mode"};
  Lowering l;
  l.walk(input);
  l.pod.sink = PGF_SINK_AGGREGATE;
  if (agg.group_expr().size() > PGF_MAX_KEYS) throw NotEligible{"too many group keys"};
  if (agg.aggr_expr().size() > PGF_MAX_AGGS) throw NotEligible{"too many aggregates"};
  int64_t groups_join = -1;
  uint32_t key_words = 0;
  std::vector<B200PipelineExec::Output> outputs;
  for (const auto& g : agg.group_expr()) {
    const ExprRef bound = l.rebase(input, g.first);  // keeps the BoundColumn alive
    auto c = bound->downcast<BoundColumn>();
    if (!c) throw NotEligible{"group key is not a column"};
    const pgf_colref ref = c->bound().ref;
    for (uint32_t j = 0; j < l.pod.njoins; ++j)
      if (l.pod.joins[j].probe_key.source == ref.source && l.pod.joins[j].probe_key.col == ref.col) groups_join = int64_t(j);
    // group keys are hashed as 64-bit words: integers one, inline views / Decimal128 two, 32 bytes in all
    const int32_t kt = c->bound().type_tag;
    const uint32_t kw = key_type_for(kt) ? 1u : (kt == PGF_T_UTF8VIEW || kt == PGF_T_BINARYVIEW || kt == PGF_T_DECIMAL128) ? 2u : 0u;
    key_words += kw;
    if (kw == 0 || key_words > 4) throw NotEligible{"group key of an unsupported type or wider than 32 bytes"};
    outputs.push_back({false, l.pod.nkeys});
    l.pod.keys[l.pod.nkeys++] = ref;
  }
  for (const auto& a : agg.aggr_expr()) {
    pgf_agg& out = l.pod.aggs[l.pod.naggs];
    const bool star = a.fun == AggregateFunction::Count &&
                      (a.args.empty() || (a.args[0]->downcast<Literal>() && a.args[0]->downcast<Literal>()->value().kind != ScalarValue::Null));
    if (star) {
      out.func = PGF_AGG_COUNT_STAR;
      out.expr = -1;
    } else {
      if (a.args.size() != 1) throw NotEligible{"aggregate with " + std::to_string(a.args.size()) + " arguments"};
      out.func = a.fun == AggregateFunction::Sum ? PGF_AGG_SUM : a.fun == AggregateFunction::Avg ? PGF_AGG_AVG : PGF_AGG_COUNT;
      out.expr = l.value_expr(l.rebase(input, a.args[0]));
    }
    outputs.push_back({true, l.pod.naggs++});
  }
  return std::make_shared<B200PipelineExec>(l.pod, lower_builds(l), l.scan, agg.schema(), outputs, groups_join);
}

// SortExec / GlobalLimitExec / column-only ProjectionExec directly above an aggregate pipeline
inline std::optional<PlanRef> absorb_above(const PlanRef& node) {
  if (node->children().size() != 1) return std::nullopt;
  auto below = node->children()[0]->downcast<B200PipelineExec>();
  if (!below || below->pod().sink != PGF_SINK_AGGREGATE) return std::nullopt;
  pgf_pipeline pod = below->pod();
  auto outputs = below->outputs();
  Schema schema = below->schema();
  if (auto s = node->downcast<SortExec>()) {
    if (pod.nsort || pod.limit || s->expr().size() > PGF_MAX_SORT) return std::nullopt;
    for (const auto& e : s->expr()) {
      auto c = e.expr->downcast<Column>();
      if (!c || c->index() >= outputs.size()) return std::nullopt;
      const auto& o = outputs[c->index()];
      pod.sort[pod.nsort++] = pgf_sort_key{o.is_agg ? 1 : 0, int32_t(o.index), e.descending ? 1 : 0, e.nulls_first ? 1 : 0};
    }
    pod.limit = s->fetch().value_or(0);
  } else if (auto g = node->downcast<GlobalLimitExec>()) {
    if (g->skip() != 0 || !g->fetch() || pod.nsort == 0 || pod.limit) return std::nullopt;
    pod.limit = *g->fetch();
  } else if (auto p = node->downcast<ProjectionExec>()) {
    std::vector<B200PipelineExec::Output> permuted;
    for (const auto& e : p->expr()) {
      auto c = e.first->downcast<Column>();
      if (!c || c->index() >= outputs.size()) return std::nullopt;
      permuted.push_back(outputs[c->index()]);
    }
    outputs = permuted;
    schema = p->schema();
  } else {
    return std::nullopt;
  }
  return PlanRef(std::make_shared<B200PipelineExec>(pod, below->builds(), below->scan(), schema, outputs, below->groups_bounded_by_join()));
}

// every pod of a rewritten tree, build sides before the pipeline that probes them
inline void collect_pods(const PlanRef& plan, std::vector<const B200PipelineExec*>& out) {
  for (const auto& c : plan->children()) collect_pods(c, out);
  if (auto p = plan->downcast<B200PipelineExec>()) out.push_back(p);
}

}  // namespace detail

// The third bottom-up rewrite (template runtime_filter_plan.rs:27-48).
// `gpu` may be null: then the library's own eligibility check (pgf_pipeline_check) is skipped and
// only the grammar decides -- used by the CPU tests of the lowering.
inline PlanRef install_b200_operators(const PlanRef& plan, const B200Context* gpu, std::vector<std::string>* skipped = nullptr) {
  std::vector<PlanRef> kids;
  for (const auto& c : plan->children()) kids.push_back(install_b200_operators(c, gpu, skipped));
  PlanRef p = kids.empty() ? plan : plan->with_new_children(std::move(kids));
  try {
    PlanRef candidate;
    if (auto agg = p->downcast<AggregateExec>()) {
      if (agg->mode() == AggregateMode::Partial) return p;  // absorbed when its Final is visited
      candidate = detail::lower_aggregate(*agg);
    } else if (auto above = detail::absorb_above(p)) {
      candidate = *above;
    } else {
      return p;
    }
    if (gpu) {
      std::vector<const B200PipelineExec*> pods;
      detail::collect_pods(candidate, pods);
      for (const B200PipelineExec* pe : pods) {
        // the library's check resolves join-table handles, which do not exist before the build
        // sides have run: pipelines that probe a join are checked when they execute
        if (pe->pod().njoins) continue;
        const pgf_status st = pgf_pipeline_check(gpu->raw(), &pe->pod());
        if (st == PGF_ERR_NOT_ELIGIBLE) {
          const char* m = pgf_last_error(gpu->raw());
          throw NotEligible{std::string("library: ") + (m ? m : "not eligible")};
        }
        gpu->check(st);
      }
    }
    return candidate;
  } catch (const NotEligible& ne) {
    if (skipped) skipped->push_back(p->name() + ": " + ne.why);
    return p;  // keep the DataFusion node
  }
}

}  // namespace pgf_b200

#endif  // PGF_B200_PLAN_HPP
