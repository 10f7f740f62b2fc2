// Explicit instantiation of the streaming pipeline kernel with the counting sink (plans with a join or a
// build sink run the compaction pipeline, probe_kernel.cuh).
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
  if (grouped || nj) return cudaErrorInvalidValue;  // build sinks and join probes never reach the streaming kernel
  return launch_one<SINK_COUNT, CLS_I64, false, 0, 1>(plan, grid, smem, stream);
}

}  // namespace pgf
