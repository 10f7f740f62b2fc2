// Explicit instantiations of the compaction pipeline (probe_kernel.cuh): one per accumulator class, generic
// predicate or the single-string-range specialisation of the TPC-H Q3 pipelines.
#include "probe_kernel.cuh"

namespace pgf {

template <uint32_t ACC, int T0, bool SPLIT = false>
static cudaError_t launch_one(const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  auto kernel = probe_pipeline_kernel<ACC, T0, SPLIT>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, kPThreads, smem, stream>>>(plan);
  return cudaGetLastError();
}

// Split execution: stages A and B (no accumulators: one instantiation serves every accumulator class), then stage C
// over the appended entries on `cgrid` CTAs.
cudaError_t launch_probe_split(uint32_t acc, int t0, const DevPlan& plan, uint32_t grid, size_t smem, uint32_t cgrid, cudaStream_t stream) {
  cudaError_t e = t0 == int(LD_VIEW) ? launch_one<CLS_F64, LD_VIEW, true>(plan, grid, smem, stream) : launch_one<CLS_F64, -1, true>(plan, grid, smem, stream);
  if (e != cudaSuccess) return e;
  switch (acc) {
    case CLS_F64: entries_pipeline_kernel<CLS_F64><<<cgrid, 256, 0, stream>>>(plan); break;
    case CLS_I64: entries_pipeline_kernel<CLS_I64><<<cgrid, 256, 0, stream>>>(plan); break;
    default: entries_pipeline_kernel<CLS_I128><<<cgrid, 256, 0, stream>>>(plan); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_probe(uint32_t acc, int t0, const DevPlan& plan, uint32_t grid, size_t smem, cudaStream_t stream) {
  const bool view1 = t0 == int(LD_VIEW);
  switch (acc) {
    case CLS_F64: return view1 ? launch_one<CLS_F64, LD_VIEW>(plan, grid, smem, stream) : launch_one<CLS_F64, -1>(plan, grid, smem, stream);
    case CLS_I64: return view1 ? launch_one<CLS_I64, LD_VIEW>(plan, grid, smem, stream) : launch_one<CLS_I64, -1>(plan, grid, smem, stream);
    default: return view1 ? launch_one<CLS_I128, LD_VIEW>(plan, grid, smem, stream) : launch_one<CLS_I128, -1>(plan, grid, smem, stream);
  }
}

cudaError_t launch_rows(uint32_t acc, const DevPlan& plan, uint32_t grid, cudaStream_t stream) {
  switch (acc) {
    case CLS_F64: rows_pipeline_kernel<CLS_F64><<<grid, 256, 0, stream>>>(plan); break;
    case CLS_I64: rows_pipeline_kernel<CLS_I64><<<grid, 256, 0, stream>>>(plan); break;
    default: rows_pipeline_kernel<CLS_I128><<<grid, 256, 0, stream>>>(plan); break;
  }
  return cudaGetLastError();
}

}  // namespace pgf
