#include "args.h"
namespace dsm {
// Givens rank-1 update of the trailing block for every deleted row (one CTA; column sweep is sequential).
__global__ void __launch_bounds__(NTHREADS) delete_rows_kernel(double* Lf, int n, const int64_t* rows, int nrows, double* v) {
  __shared__ double cs[2];
  const int tid = threadIdx.x;
  for (int q = 0; q < nrows; q++) {
    const int i = (int)rows[q] - 1;
    for (int r = i + 1 + tid; r < n; r += NTHREADS) v[r] = Lf[(int64_t)i * n + r];
    __syncthreads();
    for (int k = i + 1; k < n; k++) {
      if (tid == 0) {
        const double f = Lf[(int64_t)k * n + k], g = v[k];
        double c, s, r;
        if (g == 0.0) { c = 1.0; s = 0.0; r = f; }
        else if (f == 0.0) { c = 0.0; s = 1.0; r = g; }
        else { r = hypot(f, g); if (fabs(f) > fabs(g) && f < 0) r = -r; c = f / r; s = g / r; }
        Lf[(int64_t)k * n + k] = r; cs[0] = c; cs[1] = s;
      }
      __syncthreads();
      const double c = cs[0], s = cs[1];
      for (int r = k + 1 + tid; r < n; r += NTHREADS) {
        const double a = Lf[(int64_t)k * n + r], b = v[r];
        Lf[(int64_t)k * n + r] = c * a + s * b;
        v[r] = -s * a + c * b;
      }
      __syncthreads();
    }
  }
}

// tiled factor -> dense column-major n x n (lower triangle; upper zero)
__global__ void untile_kernel(const double* Ft, int nkc, int n, double* out) {
  const int c = blockIdx.x;
  for (int r = threadIdx.x; r < n; r += blockDim.x) out[(int64_t)c * n + r] = (r >= c) ? Ft[tidx(r, c, nkc)] : 0.0;
}
void launch_untile(const double* Ft, int nkc, int n, double* out, cudaStream_t st) { untile_kernel<<<n, 256, 0, st>>>(Ft, nkc, n, out); }
void launch_delete_rows(double* Lf, int n, const int64_t* rows, int nrows, double* v, cudaStream_t st) {
  delete_rows_kernel<<<1, NTHREADS, 0, st>>>(Lf, n, rows, nrows, v);
}
}  // namespace dsm
