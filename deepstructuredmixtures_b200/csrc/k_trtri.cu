#include "trtri.cuh"
namespace dsm {
cudaError_t init_trtri_kernels() {
  return cudaFuncSetAttribute(trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENGINE_SMEM_BYTES);
}
void launch_trtri(const TrtriArgs& a, int nctas, cudaStream_t st) { trtri_kernel<<<nctas, NTHREADS, ENGINE_SMEM_BYTES, st>>>(a); }
}  // namespace dsm
