// Registry of plan shapes with a compile-time specialised instantiation of the fused
// pipeline kernel (see "compile-time plan shapes" in pipeline_kernel.cuh).  Anything else
// runs the generic instantiation of the same kernel.
#pragma once
#include <cuda_runtime.h>

#include "device_types.cuh"

namespace pgf {

struct ShapeSig {
  uint32_t sink, acc, grouped, nj, maxe;
  int nterms;
  int term_ld[4];
  int nexprs;
  int expr_form[8];
  int nkeys;
  int key_enc[4];  // LD_* | payload << 4 | word << 8
};

using ShapeLaunchFn = cudaError_t (*)(const DevPlan&, uint32_t grid, size_t smem, cudaStream_t);
struct ShapeEntry {
  ShapeSig sig;
  ShapeLaunchFn fn;
  const char* name;
  // > 0: the instantiation takes the fast GROUP BY path (Float64 sums over NOT NULL scan columns,
  // no join) and needs shared-memory accumulator slots for this many arguments
  uint32_t fast_group_exprs;
};

const ShapeEntry* find_shape(const ShapeSig& sig);

}  // namespace pgf
