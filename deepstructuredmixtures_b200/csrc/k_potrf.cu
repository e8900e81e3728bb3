#include "chol.cuh"
namespace dsm {
void launch_solve(const SolveArgs& a, int nleaves, cudaStream_t st) { solve_kernel<<<nleaves, NTHREADS, 0, st>>>(a); }
}  // namespace dsm
