// Block back-substitution alpha = L^{-T} z for the fit-only path (gaussianprocess.jl:105); the forward solve is
// fused into the diagonal tiles of potrf2.cuh and, on the gradient path, alpha comes out of trtri3.cuh.
#pragma once
#include "common.cuh"
#include "args.h"

namespace dsm {


// Per-leaf triangular solves and reductions (one CTA per leaf):
//   z = L^{-1} y (forward), alpha = L^{-T} z (backward) with the inverse diagonal blocks W,
//   zz = z'z (= y'alpha), aa = alpha'alpha.       gaussianprocess.jl:105,163

__global__ void __launch_bounds__(NTHREADS, 1) solve_kernel(SolveArgs a) {
  __shared__ double sv[BLK];       // rhs block
  __shared__ double part[2][BLK];
  __shared__ double red[16];
  const LeafMeta m = a.meta[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkc = m.nkc;
  const double* F = a.F + m.foff;
  const double* y = a.y + m.voff;
  double* z = a.z + m.voff;
  double* al = a.alpha + m.voff;
  double zz = 0.0, aa = 0.0;
  // forward
  for (int J = 0; J < (a.skip_forward ? 0 : m.nb); J++) {
    const int w = blk_width(m.np, J), j0 = J * BLK;
    const int r = tid & (BLK - 1), half = tid >> 7;   // 2 k-halves
    double s = 0.0;
    if (r < w) {
      const int kh = (j0 / 2 + 1) & ~1;
      const int kb = half ? kh : 0, ke = half ? j0 : min(kh, j0);
      double s0 = 0, s1 = 0;
      int k = kb;
      for (; k + 1 < ke; k += 2) {
        s0 += F[tidx(j0 + r, k, nkc)] * z[k];
        s1 += F[tidx(j0 + r, k + 1, nkc)] * z[k + 1];
      }
      for (; k < ke; k++) s0 += F[tidx(j0 + r, k, nkc)] * z[k];
      s = s0 + s1;
    }
    part[half][r] = s;
    __syncthreads();
    if (tid < BLK) sv[tid] = (tid < w) ? y[j0 + tid] - (part[0][tid] + part[1][tid]) : 0.0;
    __syncthreads();
    // z_J = W_J * sv  (W lower: k <= r)
    const double* Wd = a.W + m.woff + (int64_t)J * WBLK_D;
    double t = 0.0;
    if (r < w) {
      const int kb = half ? ((r / 2 + 1) & ~1) : 0, ke = half ? r + 1 : min((r / 2 + 1) & ~1, r + 1);
      for (int k = kb; k < ke; k++) t += Wd[widx(r, k)] * sv[k];
    }
    part[half][r] = t;
    __syncthreads();
    if (tid < w) {
      const double v = part[0][tid] + part[1][tid];
      z[j0 + tid] = v;
      if (j0 + tid < m.n) zz += v * v;
    }
    __syncthreads();
  }
  // backward
  for (int J = m.nb - 1; J >= 0; J--) {
    const int w = blk_width(m.np, J), j0 = J * BLK;
    const int i1 = j0 + w;
    // s_c = sum_{i >= i1} L[i, j0+c] * alpha[i] : warp per column
    for (int c = warp; c < w; c += NTHREADS / 32) {
      double s = 0.0;
      for (int i = i1 + lane; i < m.np; i += 32) s += F[tidx(i, j0 + c, nkc)] * al[i];
      s = warp_sum(s);
      if (lane == 0) sv[c] = z[j0 + c] - s;
    }
    __syncthreads();
    // alpha_J = W_J^T sv :  alpha[c] = sum_{k >= c} W[k][c] sv[k] = sum_k WT[c + k*BLK] sv[k]
    const double* WTd = a.WT + m.woff + (int64_t)J * WBLK_D;
    const int c = tid & (BLK - 1), half = tid >> 7;
    double t = 0.0;
    if (c < w) {
      const int mid = (c + w) / 2;
      const int kb = half ? mid : c, ke = half ? w : mid;
      for (int k = kb; k < ke; k++) t += WTd[widx(c, k)] * sv[k];
    }
    part[half][c] = t;
    __syncthreads();
    if (tid < w) {
      const double v = part[0][tid] + part[1][tid];
      al[j0 + tid] = v;
      if (j0 + tid < m.n) aa += v * v;
    }
    __syncthreads();
  }
  zz = block_sum(zz, red);
  aa = block_sum(aa, red);
  if (tid == 0) { if (!a.skip_forward) a.scal[blockIdx.x].zz = zz; a.scal[blockIdx.x].aa = aa; }
}

}  // namespace dsm
