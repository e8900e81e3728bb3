#include "potrf2.cuh"
#include "trtri3.cuh"
#include "lauum3.cuh"
#include "predict3.cuh"
#include "fused2.cuh"
namespace dsm {
cudaError_t init_v2_kernels() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(potrf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES))) return e;
  if ((e = cudaFuncSetAttribute(predict3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES))) return e;
  if ((e = cudaFuncSetAttribute(lauum3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES))) return e;
  if ((e = cudaFuncSetAttribute(trtri3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES))) return e;
  if ((e = cudaFuncSetAttribute(eval2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES))) return e;
  return cudaSuccess;
}
void launch_potrf2(const Potrf2Args& a, int nctas, cudaStream_t st) { potrf2_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a); }
void launch_predict3(const PredArgs& a, int nctas, cudaStream_t st) { predict3_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a); }
void launch_predict_reduce(const PredArgs& a, cudaStream_t st) { if (a.nwcols > 0) predict_reduce_kernel<<<a.nwcols, BLK, 0, st>>>(a); }
void launch_lauum3(const LauumArgs& a, int nctas, cudaStream_t st) { lauum3_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a); }
void launch_eval2(const Potrf2Args& pa, const Trtri3Args& ta, int nctas, const int2* cols, int ncols, cudaStream_t st) {
  Eval2Args a{pa, ta};
  eval2_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a);
  alpha_reduce_kernel<<<ncols, BLK, 0, st>>>(ta, cols, ncols);
}
void launch_trtri3(const Trtri3Args& a, int nctas, const int2* cols, int ncols, cudaStream_t st) {
  trtri3_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a);
  alpha_reduce_kernel<<<ncols, BLK, 0, st>>>(a, cols, ncols);
}
void launch_eval2_only(const Potrf2Args& pa, const Trtri3Args& ta, int nctas, cudaStream_t st) {
  Eval2Args a{pa, ta};
  eval2_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a);
}
void launch_trtri3_only(const Trtri3Args& a, int nctas, cudaStream_t st) { trtri3_kernel<<<nctas, NTHREADS_PW, PIPE_SMEM_BYTES, st>>>(a); }
void launch_alpha_reduce(const Trtri3Args& a, const int2* cols, int ncols, cudaStream_t st) { alpha_reduce_kernel<<<ncols, BLK, 0, st>>>(a, cols, ncols); }
}  // namespace dsm
