#include "chol.cuh"
namespace dsm {
// every claimed task must be resident (tasks wait for each other): the grid never exceeds what fits on the device
int solve_max_ctas(int sms) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve3_kernel, NTHREADS, 0);
  return sms * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
}
void launch_solve(const SolveArgs& a, int nctas, cudaStream_t st) { solve3_kernel<<<nctas, NTHREADS, 0, st>>>(a); }
}  // namespace dsm
