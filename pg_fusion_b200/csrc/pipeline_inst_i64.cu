// Explicit instantiations of the fused pipeline kernel: AggregateExec sinks with Int64 accumulators.
#include "pipeline_kernel.cuh"

namespace pgf {

template <uint32_t SINK, uint32_t ACC, bool GROUPED, uint32_t NJ, uint32_t MAXE>
static cudaError_t launch_one(const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  auto kernel = pipeline_kernel<SINK, ACC, GROUPED, NJ, MAXE>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, pipeline_threads(SINK, GROUPED), smem, stream>>>(plan);
  return cudaGetLastError();
}

cudaError_t launch_agg_i64(bool grouped, uint32_t nj, uint32_t maxe, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  if (grouped == true && nj == 0 && maxe == 2) return launch_one<SINK_AGG, CLS_I64, true, 0, 2>(plan, grid, smem, stream);
  if (grouped == true && nj == 0 && maxe == 8) return launch_one<SINK_AGG, CLS_I64, true, 0, 8>(plan, grid, smem, stream);
  if (grouped == false && nj == 0 && maxe == 2) return launch_one<SINK_AGG, CLS_I64, false, 0, 2>(plan, grid, smem, stream);
  if (grouped == false && nj == 0 && maxe == 8) return launch_one<SINK_AGG, CLS_I64, false, 0, 8>(plan, grid, smem, stream);
  return cudaErrorInvalidValue;
}

}  // namespace pgf
