// Specialised instantiations of the fused pipeline kernel for the headline plan shapes
// (TPC-H Q6 / Q1 over the reference's Float64 + Utf8View schema).
#include "pipeline_kernel.cuh"
#include "pipeline_shapes.hpp"

namespace pgf {

namespace {

template <uint32_t SINK, uint32_t ACC, bool GROUPED, uint32_t NJ, uint32_t MAXE, class SHAPE>
cudaError_t launch_shape(const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  auto kernel = pipeline_kernel<SINK, ACC, GROUPED, NJ, MAXE, SHAPE>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, pipeline_threads(SINK, GROUPED, is_fast_grouped<SINK, ACC, GROUPED, NJ, SHAPE>(), ACC), smem, stream>>>(plan);
  return cudaGetLastError();
}

// Q6: WHERE date-range AND f64-range AND f64-range ; SUM(x * y)
using Q6Shape = ShapeT<false, IntList<LD_VIEW, LD_F64, LD_F64>, IntList<FORM_XY>, true>;
// Q1 (standard, 8 aggregates -> 5 distinct arguments) and the reference's q01.sql (7 -> 4)
// (arguments are in the canonical device order of lower_aggregate: by form; price*(1-disc)*(1+tax) reuses
// price*(1-disc) right in front of it: FORM_PREV_CPZ, the lowering checks the operands match)
using Q1Shape8 = ShapeT<false, IntList<LD_VIEW>, IntList<FORM_X, FORM_X, FORM_X, FORM_X_CMY, FORM_PREV_CPZ>, true,
                        IntList<key_enc(LD_VIEW, false, 0), key_enc(LD_VIEW, false, 2)>>;
using Q1Shape7 = ShapeT<false, IntList<LD_VIEW>, IntList<FORM_X, FORM_X, FORM_X, FORM_X_CMY>, true,
                        IntList<key_enc(LD_VIEW, false, 0), key_enc(LD_VIEW, false, 2)>>;

// "D" variants: Decimal128 money, Date32 dates, Int16 flag codes; wrapping i128 arithmetic
using Q6ShapeD = ShapeT<false, IntList<LD_I32, LD_DEC, LD_DEC>, IntList<FORM_XY>, true>;
using Q1ShapeD = ShapeT<false, IntList<LD_I32>, IntList<FORM_X, FORM_X, FORM_X, FORM_X_CMY, FORM_PREV_CPZ>, true,
                        IntList<key_enc(LD_I16, false, 0), key_enc(LD_I16, false, 1)>>;

const ShapeEntry kShapes[] = {
    {{SINK_AGG, CLS_I128, 0, 0, 2, 3, {LD_I32, LD_DEC, LD_DEC, -1}, 1, {FORM_XY, -1, -1, -1, -1, -1, -1, -1}, 0, {0, 0, 0, 0}},
     launch_shape<SINK_AGG, CLS_I128, false, 0, 2, Q6ShapeD>, "q6_decimal", 0},
    {{SINK_AGG, CLS_I128, 1, 0, kAccI128MaxExprs, 1, {LD_I32, -1, -1, -1}, 5, {FORM_X, FORM_X, FORM_X, FORM_X_CMY, FORM_PREV_CPZ, -1, -1, -1},
      2, {key_enc(LD_I16, false, 0), key_enc(LD_I16, false, 1), 0, 0}},
     launch_shape<SINK_AGG, CLS_I128, true, 0, kAccI128MaxExprs, Q1ShapeD>, "q1_decimal_8aggs", 5},
    {{SINK_AGG, CLS_F64, 0, 0, 2, 3, {LD_VIEW, LD_F64, LD_F64, -1}, 1, {FORM_XY, -1, -1, -1, -1, -1, -1, -1}, 0, {0, 0, 0, 0}},
     launch_shape<SINK_AGG, CLS_F64, false, 0, 2, Q6Shape>, "q6_f64", 0},
    {{SINK_AGG, CLS_F64, 1, 0, 8, 1, {LD_VIEW, -1, -1, -1}, 5, {FORM_X, FORM_X, FORM_X, FORM_X_CMY, FORM_PREV_CPZ, -1, -1, -1},
      2, {key_enc(LD_VIEW, false, 0), key_enc(LD_VIEW, false, 2), 0, 0}},
     launch_shape<SINK_AGG, CLS_F64, true, 0, 8, Q1Shape8>, "q1_f64_8aggs", 5},
    {{SINK_AGG, CLS_F64, 1, 0, 8, 1, {LD_VIEW, -1, -1, -1}, 4, {FORM_X, FORM_X, FORM_X, FORM_X_CMY, -1, -1, -1, -1},
      2, {key_enc(LD_VIEW, false, 0), key_enc(LD_VIEW, false, 2), 0, 0}},
     launch_shape<SINK_AGG, CLS_F64, true, 0, 8, Q1Shape7>, "q1_f64_7aggs", 4},
};

bool same(const ShapeSig& a, const ShapeSig& b) {
  if (a.sink != b.sink || a.acc != b.acc || a.grouped != b.grouped || a.nj != b.nj || a.maxe != b.maxe) return false;
  if (a.nterms != b.nterms || a.nexprs != b.nexprs) return false;
  for (int i = 0; i < a.nterms; ++i)
    if (a.term_ld[i] != b.term_ld[i]) return false;
  for (int i = 0; i < a.nexprs; ++i)
    if (a.expr_form[i] != b.expr_form[i]) return false;
  if (a.nkeys != b.nkeys) return false;
  for (int i = 0; i < a.nkeys; ++i)
    if (a.key_enc[i] != b.key_enc[i]) return false;
  return true;
}

}  // namespace

const ShapeEntry* find_shape(const ShapeSig& sig) {
  for (const ShapeEntry& e : kShapes)
    if (same(e.sig, sig)) return &e;
  return nullptr;
}

}  // namespace pgf
