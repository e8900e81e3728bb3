#include "grad.cuh"
namespace dsm {
cudaError_t init_lauum_kernels() {
  return cudaFuncSetAttribute(lauum_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENGINE_SMEM_BYTES);
}
void launch_lauum(const LauumArgs& a, int nctas, cudaStream_t st) { lauum_trace_kernel<<<nctas, NTHREADS, ENGINE_SMEM_BYTES, st>>>(a); }
void launch_rows(const RowsArgs& a, int nleaves, cudaStream_t st) { rows_kernel<<<nleaves, NTHREADS, 0, st>>>(a); }
}  // namespace dsm
