#include "predict.cuh"
namespace dsm {
cudaError_t init_predict_kernels() {
  return cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENGINE_SMEM_BYTES);
}
void launch_predict(const PredArgs& a, int nctas, cudaStream_t st) { predict_kernel<<<nctas, NTHREADS, ENGINE_SMEM_BYTES, st>>>(a); }
}  // namespace dsm
