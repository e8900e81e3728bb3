#include "chol.cuh"
namespace dsm {
cudaError_t init_potrf_kernels() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENGINE_SMEM_BYTES))) return e;
  return cudaFuncSetAttribute(potrf_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENGINE_SMEM_BYTES);
}
void launch_potrf_diag(const CholArgs& a, int nleaves, cudaStream_t st) {
  potrf_diag_kernel<<<nleaves, NTHREADS, ENGINE_SMEM_BYTES, st>>>(a);
}
void launch_potrf_panel(const CholArgs& a, int ntile_rows, int nleaves, cudaStream_t st) {
  potrf_panel_kernel<<<dim3(ntile_rows, nleaves), NTHREADS, ENGINE_SMEM_BYTES, st>>>(a);
}
void launch_solve(const SolveArgs& a, int nleaves, cudaStream_t st) { solve_kernel<<<nleaves, NTHREADS, 0, st>>>(a); }
}  // namespace dsm
