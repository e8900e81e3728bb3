// Hyper-parameter gradients  tr((alpha alpha^T - F^{-1}) dK/dtheta)  without n^3 GEMM traces.
//
// Replaces ααinvcK! (gaussianprocess.jl:219-226: n-RHS potrs), updategradients!(gp) (:165-178) and the
// kernels' updategradients! (kernels.jl:85-99,146-164,196-200,234-246: one dense n^3 GEMM per hyper).
//
//  * tr(W) and tr(W K) follow from scalars:  with F = K + cI,
//        tr(W)   = a'a - tr(F^-1),      tr(W K) = y'a - c a'a - n + c tr(F^-1),
//    and tr(F^-1) = ||L^-1||_F^2 comes out of the triangular inverse (chol.cuh).  This covers the noise
//    and sigma gradients of every kernel, IsoLinear's length scale, and the WHOLE as-written ArdSE gradient
//    (whose length-scale part is identically 0, kernels.jl:161, App. B Q3).
//  * Length-scale terms that need F^-1 element-wise (IsoSE; ArdSE/ArdLinear in mathematical mode) use
//    LAUUM tiles F^-1_IJ = sum_{K>=I} X_KI^T X_KJ on the DMMA engine with a fused epilogue that recomputes
//    dK/dlog l_h from the point tiles and reduces  sum_ij (a_i a_j - F^-1_ij) dK_ij  per tile.
#pragma once
#include "engine.cuh"
#include "args.h"

namespace dsm {


// dK_ij / dlog l_h for one pair, accumulated as  out[h] += m * dK  over the tile elements this thread owns.
__global__ void __launch_bounds__(NTHREADS, 1) lauum_trace_kernel(LauumArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task;
  __shared__ double red[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
  for (;;) {
    if (tid == 0) s_task = atomicAdd(a.counter, 1);
    __syncthreads();
    const int t = s_task;
    __syncthreads();
    if (t >= a.ntasks) return;
    const int4 tk = a.tasks[t];
    const LeafMeta m = a.meta[tk.x];
    const int I = tk.y, J = tk.z;
    const int wi = blk_width(m.np, I), wj = blk_width(m.np, J);
    const int64_t lda = m.np;
    const int nkc = m.nkc;
    const int i0 = I * BLK, j0 = J * BLK;
    const double* F = a.F + m.foff;
    const double* WTi = a.WT + m.woff + (int64_t)I * WBLK_D;
    Acc acc;
    acc_zero(acc);
    // K = I block: A[i][k] = X_II[k][i] = WT_I[i + k*BLK];  B[j][k] = X_IJ[k][j]
    if (I == J)
      mma_run<0>(acc, WTi, LDS, TILE_D, WTi, LDS, TILE_D, wi, wi, wj, false, smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
    else
      mma_run<0>(acc, WTi, LDS, TILE_D, F + tile_off(J, i0 / KC, nkc), LDS, TILE_D, wi, wi, wj, false,
                 smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
    const int k1 = i0 + wi;
    if (m.np > k1)
      mma_run<0>(acc, F + tile_off(I, k1 / KC, nkc), LDS, TILE_D, F + tile_off(J, k1 / KC, nkc), LDS, TILE_D, m.np - k1,
                 wi, wj, false, smem, 2 * CHUNK, smem + CHUNK, 2 * CHUNK);
    // ---- fused epilogue -------------------------------------------------------------------
    const double* prm = a.prm + m.poff;
    const double* x = a.xg + m.xoff;
    const double* al = a.alpha + m.voff;
    const int D = a.D, nl = m.nl, ktype = m.ktype;
    double* sxi = smem;                    // [D][BLK]
    double* sxj = smem + D * BLK;          // [D][BLK]
    double* sai = smem + 2 * D * BLK;      // [BLK]
    double* saj = sai + BLK;
    double* scf = saj + BLK;               // [D]
    for (int u = tid; u < D * BLK; u += NTHREADS) {
      const int d = u / BLK, p = u % BLK;
      sxi[u] = (p < wi) ? x[(int64_t)d * lda + i0 + p] : 0.0;
      sxj[u] = (p < wj) ? x[(int64_t)d * lda + j0 + p] : 0.0;
    }
    if (tid < BLK) { sai[tid] = (tid < wi) ? al[i0 + tid] : 0.0; saj[tid] = (tid < wj) ? al[j0 + tid] : 0.0; }
    if (tid < D) scf[tid] = (nl > 1) ? prm[PRM_COEF + tid] : prm[PRM_COEF];
    __syncthreads();
    const double v = prm[PRM_V];
    const double sym = (I == J) ? 1.0 : 2.0;
    // per-thread partials for up to 4 hypers at a time (loop over groups of hypers to bound registers)
    for (int h0 = 0; h0 < nl; h0 += 4) {
      double g[4] = {0.0, 0.0, 0.0, 0.0};
      if ((wm * 64 < wi) && (wn * 32 < wj)) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3) + e;
              if (i0 + r < m.n && j0 + c < m.n) {
                const double mij = sym * (sai[r] * saj[c] - acc[i][j][e]);
                if (ktype == ISO_SE) {
                  double r2 = 0.0;
                  for (int d = 0; d < D; d++) { const double t = sxi[d * BLK + r] - sxj[d * BLK + c]; r2 = fma(t, t, r2); }
                  const double u = scf[0] * r2;                      // -0.5 r2 / l^2
                  g[0] += mij * (v * exp(u) * (-2.0 * u));           // K * r2 / l^2
                } else if (ktype == ARD_SE) {
                  for (int hh = 0; hh < 4 && h0 + hh < nl; hh++) {
                    const int d = h0 + hh;
                    const double t = sxi[d * BLK + r] - sxj[d * BLK + c];
                    const double u = scf[d] * (t * t);
                    g[hh] += mij * (v * exp(u) * (-2.0 * u));
                  }
                } else if (ktype == ISO_LINEAR) {
                  double dot = 0.0;
                  for (int d = 0; d < D; d++) dot = fma(sxi[d * BLK + r], sxj[d * BLK + c], dot);
                  g[0] += mij * (-2.0 * scf[0] * dot);
                } else {
                  for (int hh = 0; hh < 4 && h0 + hh < nl; hh++) {
                    const int d = h0 + hh;
                    g[hh] += mij * (-2.0 * scf[d] * sxi[d * BLK + r] * sxj[d * BLK + c]);
                  }
                }
              }
            }
      }
      for (int hh = 0; hh < 4 && h0 + hh < nl; hh++) {
        const double s = block_sum(g[hh], red);
        if (tid == 0) a.gpart[a.gpart_off[tk.x] + (int64_t)tk.w * nl + h0 + hh] = s;
      }
    }
    __syncthreads();
  }
}

// Per-leaf rows [lml, g_l(1..nl), g_sigma, g_noise, 0...]  (gaussianprocess.jl:163,176,206-217).
__global__ void rows_kernel(RowsArgs a) {
  __shared__ double red[16];
  const int slot = blockIdx.x, tid = threadIdx.x;
  const LeafMeta m = a.meta[slot];
  LeafScal sc = a.scal[slot];
  const double* prm = a.prm + m.poff;
  double* row = a.rows + (int64_t)m.leaf * a.row_width;
  const double n = (double)m.n;
  if (a.ldpart != nullptr) {      // engine v2: ordered sums of the per-block-column partials
    double l = 0.0, zq = 0.0;
    for (int i = tid; i < m.nb; i += blockDim.x) { l += a.ldpart[a.trpart_off[slot] / 2 + i]; zq += a.zzpart[a.trpart_off[slot] / 2 + i]; }
    sc.logdet = block_sum(l, red);
    sc.zz = block_sum(zq, red);
  }
  if (a.alpha != nullptr && a.with_grad && !(a.mask != nullptr && a.mask[slot] == 0)) {
    double q = 0.0;
    for (int i = tid; i < m.n; i += blockDim.x) { const double v = a.alpha[m.voff + i]; q = fma(v, v, q); }
    sc.aa = block_sum(q, red);
  }
  double lml = -(sc.zz + sc.logdet + 1.8378770664093453 * n) / 2.0;   // log(2 pi)
  if (sc.info != 0) lml = nan("");
  if (!a.with_grad || (a.mask != nullptr && a.mask[slot] == 0)) {
    if (tid == 0) {
      row[0] = lml;
      if (a.with_grad) for (int h = 1; h < a.row_width; h++) row[h] = 0.0;      // skipped expert: weight 0 in the down-pass
    }
    return;
  }
  // tr(F^-1): ordered sum of the per-block partials
  double tr = 0.0;
  for (int i = tid; i < 2 * m.nb; i += blockDim.x) tr += a.trpart[a.trpart_off[slot] + i];
  tr = block_sum(tr, red);
  const double c = prm[PRM_C], eta = prm[PRM_ETA], s = prm[PRM_S];
  const double trW = sc.aa - tr;
  const double trWK = sc.zz - c * sc.aa - n + c * tr;
  const bool se = (m.ktype == ISO_SE || m.ktype == ARD_SE);
  const double sfac = (a.as_written && se) ? s : 1.0;
  const int ntask = m.nb * (m.nb + 1) / 2;
  for (int h = 0; h < m.nl; h++) {
    double gl = 0.0;
    const bool shortcut = (m.ktype == ISO_LINEAR) || (m.ktype == ARD_SE && a.as_written);
    if (m.ktype == ISO_LINEAR) {
      gl = -trWK;                                    // kernels.jl:198
    } else if (m.ktype == ARD_SE && a.as_written) {
      gl = 0.0;                                      // kernels.jl:161 (App. B Q3)
    } else if (a.lauum_ran) {
      double p = 0.0;
      for (int i = tid; i < ntask; i += blockDim.x) p += a.gpart[a.gpart_off[slot] + (int64_t)i * m.nl + h];
      p = block_sum(p, red);
      gl = 0.5 * sfac * p;
    }
    (void)shortcut;
    if (tid == 0) row[1 + h] = gl;
  }
  if (tid == 0) {
    row[0] = lml;
    row[1 + m.nl] = se ? sfac * trWK : 0.0;          // kernels.jl:93,157 ; linear: 0 (kernels.jl:201)
    row[2 + m.nl] = eta * trW;                       // gaussianprocess.jl:176
    for (int h = 3 + m.nl; h < a.row_width; h++) row[h] = 0.0;
    a.scal[slot].trinv = tr;
  }
}

}  // namespace dsm
