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
//    dK/dlog l_h from the point tiles and reduces  sum_ij (a_i a_j - F^-1_ij) dK_ij  per tile (lauum3.cuh).
#pragma once
#include "common.cuh"
#include "args.h"

namespace dsm {


// Per-leaf rows [lml, g_l(1..nl), g_sigma, g_noise, 0...]  (gaussianprocess.jl:163,176,206-217).
__global__ void rows_kernel(RowsArgs a) {
  __shared__ double red[16];
  const int slot = blockIdx.x, tid = threadIdx.x;
  if (a.share != nullptr && a.share[slot].x == SHARE_ALIAS) return;       // row copied from the source by rows_alias_kernel
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
