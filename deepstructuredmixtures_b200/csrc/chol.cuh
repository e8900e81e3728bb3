// Batched variable-size blocked FP64 Cholesky / triangular inverse on the DMMA engine.
//
// Replaces LAPACK.potrf!('L', F) (gaussianprocess.jl:101, AdvancedCholeskey.jl:171), the two
// triangular solves of gaussianprocess.jl:105 and the n-RHS ldiv!(cK, -I) of :219-226.
//
// Left-looking by 128-wide block columns.  For block column J of every expert:
//   diag task   : C = F_JJ - sum_{K<J} L_JK L_JK^T ; L_JJ = chol(C) ; W_J = L_JJ^{-1} (kept, plus W_J^T)
//   panel task  : C = F_IJ - sum_{K<J} L_IK L_JK^T ; L_IJ = C W_J^T            (I > J)
// A block column step is one launch of each task type over ALL experts that still have that column
// (level-synchronous schedule).  The triangular inverse X = L^{-1} needs no cross-task ordering at
// all: block column J of X depends only on L and on itself.
#pragma once
#include "engine.cuh"
#include "args.h"

namespace dsm {


// grid.x = number of leaves (slot order), one CTA per leaf; CTAs of leaves without block column `step` exit.
__global__ void __launch_bounds__(NTHREADS, 1) potrf_diag_kernel(CholArgs a) {
  extern __shared__ __align__(16) double smem[];
  const LeafMeta m = a.meta[blockIdx.x];
  const int J = a.step;
  if (J >= m.nb) return;
  const int tid = threadIdx.x;
  const int w = blk_width(m.np, J);
  const int64_t lda = m.np;
  const int j0 = J * BLK;
  double* F = a.F + m.foff;
  double* S = smem;                 // resident tile (aliases REGION0)
  double* aux = smem + REGION0;     // REGION1: scratch (col buffer, flags)
  const bool factor = J >= a.jstart;

  if (factor) {
    Acc acc;
    acc_zero(acc);
    if (j0 > 0)
      mma_run<0>(acc, F + j0, lda, F + j0, lda, j0, w, w, true, smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
    acc_store_colmajor(acc, S, 1.0);
    __syncthreads();
  }
  // C = F_JJ - acc (lower part)
  for (int c = tid >> 5; c < w; c += NTHREADS / 32) {
    const double* src = F + (int64_t)(j0 + c) * lda + j0;
    for (int r = (tid & 31); r < w; r += 32) {
      if (r >= c) S[c * LDS + r] = factor ? src[r] - S[c * LDS + r] : src[r];
      else S[c * LDS + r] = 0.0;
    }
  }
  if (tid == 0) { aux[0] = 0.0; aux[1] = 0.0; }
  __syncthreads();
  if (factor) {
    const int info = potrf_smem(S, w, aux);
    if (tid == 0 && info != 0 && a.scal[blockIdx.x].info == 0) a.scal[blockIdx.x].info = j0 + info;
    // write L_JJ (lower) back; logdet partial
    double ld = 0.0;
    for (int c = tid >> 5; c < w; c += NTHREADS / 32) {
      double* dst = F + (int64_t)(j0 + c) * lda + j0;
      for (int r = (tid & 31); r < w; r += 32)
        if (r >= c) dst[r] = S[c * LDS + r];
    }
    for (int r = tid; r < w; r += NTHREADS)
      if (j0 + r < m.n) ld += log(S[r * LDS + r]);
    ld = block_sum(ld, aux + 8);
    if (tid == 0) {
      if (J == 0) a.scal[blockIdx.x].logdet = 2.0 * ld;
      else a.scal[blockIdx.x].logdet += 2.0 * ld;
    }
  }
  __syncthreads();
  trtri_smem(S, w, aux + 16);
  // W (column-major, upper zero) and W^T; tr(F^{-1}) diagonal-block partial over the real rows/cols
  double* Wd = a.W + m.woff + (int64_t)J * BLK * BLK;
  double* WTd = a.WT + m.woff + (int64_t)J * BLK * BLK;
  double tr = 0.0;
  for (int c = tid >> 5; c < BLK; c += NTHREADS / 32) {
    for (int r = (tid & 31); r < BLK; r += 32) {
      double v = 0.0;
      if (r < w && c < w && r >= c) v = S[c * LDS + r];
      Wd[c * BLK + r] = v;
      if (j0 + r < m.n && j0 + c < m.n) tr += v * v;
    }
  }
  // transposed copy: WT[c + r*BLK] = W[r][c]  -> iterate with c fastest for coalescing
  for (int r = tid >> 5; r < BLK; r += NTHREADS / 32) {
    for (int c = (tid & 31); c < BLK; c += 32) {
      double v = 0.0;
      if (r < w && c < w && r >= c) v = S[c * LDS + r];
      WTd[r * BLK + c] = v;
    }
  }
  tr = block_sum(tr, aux + 8);
  if (tid == 0) a.trpart[a.trpart_off[blockIdx.x] + J] = tr;
}

// grid = (max_nb - step - 1, leaves); CTA (x, y): tile row I = step + 1 + x of leaf y.
__global__ void __launch_bounds__(NTHREADS, 1) potrf_panel_kernel(CholArgs a) {
  extern __shared__ __align__(16) double smem[];
  const LeafMeta m = a.meta[blockIdx.y];
  const int J = a.step, I = J + 1 + blockIdx.x;
  if (I >= m.nb) return;
  if (I < a.jstart) return;          // chol_continue: already final
  const int tid = threadIdx.x;
  const int wi = blk_width(m.np, I), wj = blk_width(m.np, J);
  const int64_t lda = m.np;
  const int i0 = I * BLK, j0 = J * BLK;
  double* F = a.F + m.foff;
  double* S = smem;
  Acc acc;
  acc_zero(acc);
  if (j0 > 0)
    mma_run<0>(acc, F + i0, lda, F + j0, lda, j0, wi, wj, false, smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
  acc_store_colmajor(acc, S, 1.0);
  __syncthreads();
  for (int c = tid >> 5; c < wj; c += NTHREADS / 32) {
    const double* src = F + (int64_t)(j0 + c) * lda + i0;
    for (int r = (tid & 31); r < wi; r += 32) S[c * LDS + r] = src[r] - S[c * LDS + r];
  }
  __syncthreads();
  // X = C * W_J^T :  X[r][c] = sum_k C[r][k] W[c][k]  (A resident, B = W streamed; W lower => k <= c)
  acc_zero(acc);
  const double* Wd = a.W + m.woff + (int64_t)J * BLK * BLK;
  mma_run<1>(acc, nullptr, 0, Wd, BLK, wj, wi, wj, false, S, 0, smem + REGION0, CHUNK, false, true);
  acc_store_colmajor(acc, S, 1.0);
  __syncthreads();
  for (int c = tid >> 5; c < wj; c += NTHREADS / 32) {
    double* dst = F + (int64_t)(j0 + c) * lda + i0;
    for (int r = (tid & 31); r < wi; r += 32) dst[r] = S[c * LDS + r];
  }
}

// Per-leaf triangular solves and reductions (one CTA per leaf):
//   z = L^{-1} y (forward), alpha = L^{-T} z (backward) with the inverse diagonal blocks W,
//   zz = z'z (= y'alpha), aa = alpha'alpha.       gaussianprocess.jl:105,163

__global__ void __launch_bounds__(NTHREADS, 1) solve_kernel(SolveArgs a) {
  __shared__ double sv[BLK];       // rhs block
  __shared__ double part[2][BLK];
  __shared__ double red[16];
  const LeafMeta m = a.meta[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t lda = m.np;
  const double* F = a.F + m.foff;
  const double* y = a.y + m.voff;
  double* z = a.z + m.voff;
  double* al = a.alpha + m.voff;
  double zz = 0.0, aa = 0.0;
  // forward
  for (int J = 0; J < m.nb; J++) {
    const int w = blk_width(m.np, J), j0 = J * BLK;
    const int r = tid & (BLK - 1), half = tid >> 7;   // 2 k-halves
    double s = 0.0;
    if (r < w) {
      const int kh = (j0 / 2 + 1) & ~1;
      const int kb = half ? kh : 0, ke = half ? j0 : min(kh, j0);
      const double* row = F + j0 + r;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int k = kb;
      for (; k + 3 < ke; k += 4) {
        s0 += row[(int64_t)k * lda] * z[k];
        s1 += row[(int64_t)(k + 1) * lda] * z[k + 1];
        s2 += row[(int64_t)(k + 2) * lda] * z[k + 2];
        s3 += row[(int64_t)(k + 3) * lda] * z[k + 3];
      }
      for (; k < ke; k++) s0 += row[(int64_t)k * lda] * z[k];
      s = (s0 + s1) + (s2 + s3);
    }
    part[half][r] = s;
    __syncthreads();
    if (tid < BLK) sv[tid] = (tid < w) ? y[j0 + tid] - (part[0][tid] + part[1][tid]) : 0.0;
    __syncthreads();
    // z_J = W_J * sv  (W lower: k <= r)
    const double* Wd = a.W + m.woff + (int64_t)J * BLK * BLK;
    double t = 0.0;
    if (r < w) {
      const int kb = half ? ((r / 2 + 1) & ~1) : 0, ke = half ? r + 1 : min((r / 2 + 1) & ~1, r + 1);
      for (int k = kb; k < ke; k++) t += Wd[k * BLK + r] * sv[k];
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
      const double* colp = F + (int64_t)(j0 + c) * lda;
      double s = 0.0;
      for (int i = i1 + lane; i < m.np; i += 32) s += colp[i] * al[i];
      s = warp_sum(s);
      if (lane == 0) sv[c] = z[j0 + c] - s;
    }
    __syncthreads();
    // alpha_J = W_J^T sv :  alpha[c] = sum_{k >= c} W[k][c] sv[k] = sum_k WT[c + k*BLK] sv[k]
    const double* WTd = a.WT + m.woff + (int64_t)J * BLK * BLK;
    const int c = tid & (BLK - 1), half = tid >> 7;
    double t = 0.0;
    if (c < w) {
      const int mid = (c + w) / 2;
      const int kb = half ? mid : c, ke = half ? w : mid;
      for (int k = kb; k < ke; k++) t += WTd[k * BLK + c] * sv[k];
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
  if (tid == 0) { a.scal[blockIdx.x].zz = zz; a.scal[blockIdx.x].aa = aa; }
}

}  // namespace dsm
