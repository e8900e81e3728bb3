// Batched triangular inverse X = L^{-1} on the DMMA engine (the n-RHS ldiv!(cK, -I) of gaussianprocess.jl:219-226
// reduced to what the gradients need: ||X||_F^2 = tr(F^{-1}), and X itself for the LAUUM pass).
#pragma once
#include "engine.cuh"
#include "args.h"

namespace dsm {

// Triangular inverse X = L^{-1}, stored TRANSPOSED in the strict upper block triangle of the factor
// (X_IJ^T at rows of block J, columns of block I), diagonal blocks in W/WT.  Task = (leaf, block column J),
// tasks[] sorted by decreasing cost; CTAs claim tasks through an atomic counter.
// Also accumulates tr(F^{-1}) = ||X||_F^2 partials per task.

__global__ void __launch_bounds__(NTHREADS, 1) trtri_kernel(TrtriArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task;
  __shared__ double red[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
  double* S = smem;
  for (;;) {
    if (tid == 0) s_task = atomicAdd(a.counter, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= a.ntasks) return;
    const int2 tk = a.tasks[t];
    const LeafMeta m = a.meta[tk.x];
    const int J = tk.y;
    const int wj = blk_width(m.np, J);
    const int64_t lda = m.np;
    const int j0 = J * BLK;
    double* F = a.F + m.foff;
    const double* WTj = a.WT + m.woff + (int64_t)J * BLK * BLK;
    double tr = 0.0;
    for (int I = J + 1; I < m.nb; I++) {
      const int wi = blk_width(m.np, I);
      const int i0 = I * BLK;
      Acc acc;
      acc_zero(acc);
      // S = sum_{K=J}^{I-1} L_IK X_KJ ;  K = J block: X_JJ = W_J  ->  B[j][k] = W_J[k][j] = WT_J[j + k*BLK]
      mma_run<0>(acc, F + i0 + (int64_t)j0 * lda, lda, WTj, BLK, wj, wi, wj, false,
                 smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
      const int k1 = j0 + wj;
      if (i0 > k1)
        mma_run<0>(acc, F + i0 + (int64_t)k1 * lda, lda, F + j0 + (int64_t)k1 * lda, lda, i0 - k1, wi, wj, false,
                   smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
      // T = -W_I * S :  T[r][c] = -sum_k W_I[r][k] S[k][c]   (A = W_I streamed, lower => k <= r; B resident)
      acc_store_rowmajor(acc, S, 1.0);
      __syncthreads();
      acc_zero(acc);
      const double* Wi = a.W + m.woff + (int64_t)I * BLK * BLK;
      mma_run<2>(acc, Wi, BLK, nullptr, 0, wi, wi, wj, false, smem + REGION0, CHUNK, S, 0, true, false);
      // write X_IJ^T into the upper triangle: element (r, c) -> F[(j0 + c) + (i0 + r) * lda]
      if ((wm * 64 < wi) && (wn * 32 < wj)) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3);
            const double v0 = -acc[i][j][0], v1 = -acc[i][j][1];
            *reinterpret_cast<double2*>(F + (int64_t)(i0 + r) * lda + j0 + c) = make_double2(v0, v1);
            if (i0 + r < m.n) {
              if (j0 + c < m.n) tr += v0 * v0;
              if (j0 + c + 1 < m.n) tr += v1 * v1;
            }
          }
      }
      __syncthreads();   // global writes of this block visible to the CTA's next cp.async reads
    }
    tr = block_sum(tr, red);
    if (tid == 0) a.trpart[a.trpart_off[tk.x] + m.nb + J] = tr;   // second half of the leaf's partial array
  }
}


}  // namespace dsm
