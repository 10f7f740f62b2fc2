// Explicit instantiations of the fused pipeline kernel: AggregateExec sinks with Decimal128 (i128) accumulators.
#include "pipeline_kernel.cuh"

namespace pgf {

template <uint32_t SINK, uint32_t ACC, bool GROUPED, uint32_t NJ>
static cudaError_t launch_one(const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  auto kernel = pipeline_kernel<SINK, ACC, GROUPED, NJ>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, pipeline_threads(SINK, GROUPED), smem, stream>>>(plan);
  return cudaGetLastError();
}

cudaError_t launch_agg_i128(bool grouped, uint32_t nj, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  if (grouped) return nj ? launch_one<SINK_AGG, CLS_I128, true, 1>(plan, grid, smem, stream)
                          : launch_one<SINK_AGG, CLS_I128, true, 0>(plan, grid, smem, stream);
  return nj ? launch_one<SINK_AGG, CLS_I128, false, 1>(plan, grid, smem, stream)
            : launch_one<SINK_AGG, CLS_I128, false, 0>(plan, grid, smem, stream);
}

}  // namespace pgf
