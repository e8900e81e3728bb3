#include "gram.cuh"
namespace dsm {
void launch_gram_fit(const GramArgs& a, int64_t ntiles, cudaStream_t st) {
  gram_fit_kernel<<<(unsigned)ntiles, NTHREADS, 0, st>>>(a);
}
void launch_gram_rect(const GramRectArgs& a, cudaStream_t st) {
  dim3 grid((unsigned)((a.na + GT - 1) / GT), (unsigned)((a.nb + GT - 1) / GT));
  gram_rect_kernel<<<grid, NTHREADS, 0, st>>>(a);
}
void launch_gather(const GatherArgs& a, int maxnp, int nleaves, cudaStream_t st) {
  dim3 grid((maxnp + 255) / 256, nleaves);
  gather_kernel<<<grid, 256, 0, st>>>(a);
}
}  // namespace dsm
