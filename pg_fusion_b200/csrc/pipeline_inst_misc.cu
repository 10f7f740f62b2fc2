// Explicit instantiations of the fused pipeline kernel: HashJoinExec build sink and the counting sink.
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

cudaError_t launch_build_or_count(bool grouped, uint32_t nj, uint32_t maxe, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  (void)maxe;
  if (grouped)  // "grouped" selects the join-build sink here
    return nj ? launch_one<SINK_JOIN_BUILD, CLS_I64, false, 1, 1>(plan, grid, smem, stream)
              : launch_one<SINK_JOIN_BUILD, CLS_I64, false, 0, 1>(plan, grid, smem, stream);
  return nj ? launch_one<SINK_COUNT, CLS_I64, false, 1, 1>(plan, grid, smem, stream)
            : launch_one<SINK_COUNT, CLS_I64, false, 0, 1>(plan, grid, smem, stream);
}

}  // namespace pgf
