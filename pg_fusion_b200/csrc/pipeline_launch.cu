// Dispatch from the runtime plan shape to the matching kernel instantiation.
#include "pipeline_kernel.cuh"

namespace pgf {

cudaError_t launch_agg_f64(bool grouped, uint32_t nj, uint32_t maxe, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream);
cudaError_t launch_agg_i64(bool grouped, uint32_t nj, uint32_t maxe, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream);
cudaError_t launch_agg_i128(bool grouped, uint32_t nj, uint32_t maxe, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream);
cudaError_t launch_build_or_count(bool build, uint32_t nj, uint32_t maxe, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream);

cudaError_t launch_pipeline(uint32_t sink, uint32_t acc, bool grouped, uint32_t nj, uint32_t maxe, const DevPlan& plan,
                            uint32_t grid, size_t smem, cudaStream_t stream) {
  if (sink == SINK_AGG) {
    switch (acc) {
      case CLS_F64: return launch_agg_f64(grouped, nj, maxe, plan, grid, smem, stream);
      case CLS_I64: return launch_agg_i64(grouped, nj, maxe, plan, grid, smem, stream);
      default: return launch_agg_i128(grouped, nj, maxe, plan, grid, smem, stream);
    }
  }
  return launch_build_or_count(sink == SINK_JOIN_BUILD, nj, maxe, plan, grid, smem, stream);
}

}  // namespace pgf
