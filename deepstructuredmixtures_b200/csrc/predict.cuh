// Batched predictive mean / variance of the leaf experts.
//
// Replaces prediction(gp, xtest) gaussianprocess.jl:110-137:  mu = m + Knt' alpha ;  V = L \ Knt ;
// Sigma = Ktt - V'V ; diag += exp(2 logNoise).  Only diag(Sigma) is consumed (common.jl:136,147), so the
// T x T matrices Ktt and V'V are never formed:  var_t = k(x_t,x_t) - sum_k V_kt^2 + eta.
//
// Task = (leaf, block Q of 128 routed test points).  Block forward substitution on the DMMA engine:
//   for I = 0..nb-1:  S = Knt_IQ - sum_{K<I} L_IK V_KQ ;  V_IQ = W_I S      (W_I = L_II^{-1} from the fit)
// with Knt_IQ recomputed from the point tiles in the epilogue (never stored) and V kept transposed in a
// per-leaf scratch so that both engine operands are contiguous along the tile dimension.
#pragma once
#include "engine.cuh"
#include "args.h"

namespace dsm {



__device__ __forceinline__ double kernel_pair(int ktype, int D, const double* sxa, int ra, const double* sxb, int rb,
                                              const double* scf, double v) {
  if (ktype == ISO_SE) {
    double r2 = 0.0;
    for (int d = 0; d < D; d++) { const double t = sxa[d * BLK + ra] - sxb[d * BLK + rb]; r2 = fma(t, t, r2); }
    return v * exp(scf[0] * r2);
  } else if (ktype == ARD_SE) {
    double s = 0.0;
    for (int d = 0; d < D; d++) { const double t = sxa[d * BLK + ra] - sxb[d * BLK + rb]; s += exp(scf[d] * (t * t)); }
    return v * s;
  } else if (ktype == ISO_LINEAR) {
    double s = 0.0;
    for (int d = 0; d < D; d++) s = fma(sxa[d * BLK + ra], sxb[d * BLK + rb], s);
    return scf[0] * s;
  } else {
    double s = 0.0;
    for (int d = 0; d < D; d++) s = fma(scf[d] * sxa[d * BLK + ra], sxb[d * BLK + rb], s);
    return s;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) predict_kernel(PredArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task;
  __shared__ double s_sq[2][BLK], s_mu[2][BLK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
  double* S = smem;
  // epilogue staging lives in REGION1 (REGION0 is the resident tile / pipeline)
  for (;;) {
    if (tid == 0) s_task = atomicAdd(a.counter, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= a.ntasks) return;
    const int2 tk = a.tasks[t];
    const PredLeaf pl = a.pl[tk.x];
    const LeafMeta m = a.meta[pl.slot];
    const int Q = tk.y, q0 = Q * BLK;
    const int wq = min(BLK, pl.Tp - q0);
    const int64_t lda = m.np, ldv = pl.Tp;
    const double* F = a.F + m.foff;
    const double* x = a.xg + m.xoff;
    const double* xt = a.xt + pl.xtoff;
    const double* al = a.alpha + m.voff;
    const double* prm = a.prm + m.poff;
    double* VT = a.VT + pl.vtoff;
    const int D = a.D, ktype = m.ktype;
    const double v = prm[PRM_V];
    if (tid < BLK) { s_sq[0][tid] = 0.0; s_sq[1][tid] = 0.0; s_mu[0][tid] = 0.0; s_mu[1][tid] = 0.0; }
    double* sxi = smem + REGION0;              // [D][BLK] rows of block I
    double* sxq = sxi + D * BLK;               // [D][BLK] test points of block Q
    double* sal = sxq + D * BLK;               // [BLK] alpha of block I
    double* scf = sal + BLK;                   // [D]
    for (int I = 0; I < m.nb; I++) {
      const int wi = blk_width(m.np, I), i0 = I * BLK;
      Acc acc;
      acc_zero(acc);
      if (i0 > 0)
        mma_run<0>(acc, F + tile_off(I, 0, m.nkc), LDS, TILE_D, VT + q0, ldv, (int64_t)KC * ldv, i0, wi, wq, false,
                   smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
      // stage point tiles (REGION1 is idle between engine runs)
      for (int u = tid; u < D * BLK; u += NTHREADS) {
        const int d = u / BLK, p = u % BLK;
        sxi[u] = (p < wi) ? x[(int64_t)d * lda + i0 + p] : 0.0;
        sxq[u] = (p < wq) ? xt[(int64_t)d * ldv + q0 + p] : 0.0;
      }
      if (tid < BLK) sal[tid] = (tid < wi && i0 + tid < m.n) ? al[i0 + tid] : 0.0;
      if (tid < D) scf[tid] = (m.nl > 1) ? prm[PRM_COEF + tid] : prm[PRM_COEF];
      __syncthreads();
      // S = Knt_IQ - acc ; mean partial sum_r Knt[r][t] alpha[r]
      double mpart[8];
#pragma unroll
      for (int q = 0; q < 8; q++) mpart[q] = 0.0;
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3) + e;
            double k = 0.0;
            if (i0 + r < m.n && q0 + c < pl.T) k = kernel_pair(ktype, D, sxi, r, sxq, c, scf, v);
            mpart[j * 2 + e] = fma(k, sal[r], mpart[j * 2 + e]);
            acc[i][j][e] = k - acc[i][j][e];
          }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        double s = mpart[q];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if ((lane >> 2) == 0) s_mu[wm][wn * 32 + (q >> 1) * 8 + 2 * (lane & 3) + (q & 1)] += s;
      }
      acc_store_rowmajor(acc, S, 1.0);
      __syncthreads();
      acc_zero(acc);
      const double* Wi = a.W + m.woff + (int64_t)I * WBLK_D;
      mma_run<2>(acc, Wi, LDS, TILE_D, nullptr, 0, 0, wi, wi, wq, false, smem + REGION0, CHUNK, S, 0, true, false);
      // V_IQ -> VT[(q0 + c) + (i0 + r) * ldv] ; column sums of squares
      double spart[8];
#pragma unroll
      for (int q = 0; q < 8; q++) spart[q] = 0.0;
      if ((wm * 64 < wi) && (wn * 32 < wq)) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3);
            const double v0 = acc[i][j][0], v1 = acc[i][j][1];
            *reinterpret_cast<double2*>(VT + (int64_t)(i0 + r) * ldv + q0 + c) = make_double2(v0, v1);
            spart[j * 2] = fma(v0, v0, spart[j * 2]);
            spart[j * 2 + 1] = fma(v1, v1, spart[j * 2 + 1]);
          }
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {
        double s = spart[q];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if ((lane >> 2) == 0) s_sq[wm][wn * 32 + (q >> 1) * 8 + 2 * (lane & 3) + (q & 1)] += s;
      }
      __syncthreads();
    }
    // finish: mu = m + Knt' alpha ; var = k(x_t, x_t) - sum V^2 + eta      (gaussianprocess.jl:117-126)
    if (tid < wq && q0 + tid < pl.T) {
      double ktt;
      if (ktype == ISO_SE) ktt = v;
      else if (ktype == ARD_SE) ktt = v * (double)D;
      else {
        ktt = 0.0;
        for (int d = 0; d < D; d++) {
          const double xv = xt[(int64_t)d * ldv + q0 + tid];
          const double cf = (ktype == ISO_LINEAR) ? prm[PRM_COEF] : prm[PRM_COEF + d];
          ktt = fma(cf * xv, xv, ktt);
        }
      }
      a.mu[pl.ooff + q0 + tid] = a.leaf_mean[m.leaf] + (s_mu[0][tid] + s_mu[1][tid]);
      a.var[pl.ooff + q0 + tid] = ktt - (s_sq[0][tid] + s_sq[1][tid]) + prm[PRM_ETA];
    }
    __syncthreads();
  }
}

}  // namespace dsm
