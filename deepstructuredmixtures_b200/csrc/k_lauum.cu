#include "grad.cuh"
namespace dsm {
void launch_rows(const RowsArgs& a, int nleaves, cudaStream_t st) { rows_kernel<<<nleaves, NTHREADS, 0, st>>>(a); }
}  // namespace dsm
