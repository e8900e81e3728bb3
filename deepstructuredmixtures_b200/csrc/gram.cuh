// Gram-matrix construction for IsoSE / ArdSE / IsoLinear / ArdLinear.
//
// Replaces getdistancematrix + kernelmatrix! (kernels.jl:15-53,78-83,133-144,189-194,228-232) and the
// noise add of update_cholesky! (gaussianprocess.jl:90-98).  The distance tensor P (n x n x D in the
// reference, stored for the model's lifetime) is never materialised: each 64 x 64 output tile stages
// the two point tiles in shared memory (SoA by dimension) and recomputes the differences in registers.
//   IsoSE     K = v exp(-0.5 |xi-xj|^2 / l^2)                   kernels.jl:21-26,78,83
//   ArdSE     K = v sum_d exp(-0.5 (xi_d-xj_d)^2 / l_d^2)       kernels.jl:31-49 (ADDITIVE over d, in order d=1..D)
//   IsoLinear K = (xi . xj) / l^2                               kernels.jl:189,194
//   ArdLinear K = sum_d xi_d xj_d / l_d^2                       SURVEY App. A.2 (reference non-functional)
#pragma once
#include "common.cuh"
#include "fastexp.cuh"
#include "args.h"

namespace dsm {


// Computes a GT x GT tile: rows ra0.. of point set A, cols rb0.. of point set B.
// xa / xb: SoA inputs, dimension d of point p at xa[d*sa + p].  Thread (tx, ty) owns rows tx*4..+3, cols ty*4..+3.
__device__ __forceinline__ void gram_tile(int ktype, int D, const double* __restrict__ prm,
                                          const double* __restrict__ xa, int64_t sa, int ra0, int na,
                                          const double* __restrict__ xb, int64_t sb, int rb0, int nb,
                                          double (&out)[4][4], double* sxa, double* sxb, double* scoef, const double* sT) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  const bool ard = (ktype == ARD_SE || ktype == ARD_LINEAR);
  for (int d0 = 0; d0 < D; d0 += GDC) {
    const int dc = min(GDC, D - d0);
    __syncthreads();
    for (int u = tid; u < dc * GT; u += NTHREADS) {
      const int d = u / GT, p = u % GT;
      sxa[d * GT + p] = (ra0 + p < na) ? xa[(int64_t)(d0 + d) * sa + ra0 + p] : 0.0;
      sxb[d * GT + p] = (rb0 + p < nb) ? xb[(int64_t)(d0 + d) * sb + rb0 + p] : 0.0;
    }
    if (tid < dc) scoef[tid] = ard ? prm[PRM_COEF + d0 + tid] : prm[PRM_COEF];
    __syncthreads();
    for (int d = 0; d < dc; d++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sxa[d * GT + tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sxb[d * GT + ty * 4 + j];
      const double cf = scoef[d];
      if (ktype == ISO_SE) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) { const double t = a[i] - b[j]; acc[i][j] = fma(t, t, acc[i][j]); }
      } else if (ktype == ARD_SE) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) { const double t = a[i] - b[j]; acc[i][j] += exp_neg(cf * (t * t), sT); }
      } else if (ktype == ISO_LINEAR) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fma(cf * a[i], b[j], acc[i][j]);
      }
    }
  }
  const double v = prm[PRM_V], c0 = prm[PRM_COEF];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      double k;
      if (ktype == ISO_SE) k = v * exp_neg(c0 * acc[i][j], sT);
      else if (ktype == ARD_SE) k = v * acc[i][j];
      else if (ktype == ISO_LINEAR) k = c0 * acc[i][j];
      else k = acc[i][j];
      out[i][j] = k;
    }
}


// F = K + (eta + 1e-8) I on the lower triangle (diagonal tiles written in full), identity on the padding.
__global__ void __launch_bounds__(NTHREADS) gram_fit_kernel(GramArgs a) {
  __shared__ double sxa[GDC * GT], sxb[GDC * GT], scoef[GDC], sT[EXPTAB_N];
  exptab_load(sT);                 // visible after the first barrier inside gram_tile
  // leaf of this tile: binary search in tile_off
  const int64_t g = blockIdx.x;
  int lo = 0, hi = a.nleaves;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.tile_off[mid] <= g) lo = mid; else hi = mid; }
  const LeafMeta m = a.meta[lo];
  const int t = (int)(g - a.tile_off[lo]);
  // t = ti*(ti+1)/2 + tj, tj <= ti
  int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
  while (ti * (ti + 1) / 2 > t) ti--;
  const int tj = t - ti * (ti + 1) / 2;
  if (a.share != nullptr) {          // block-uniform: aliased experts are not built, copied block rows neither
    const int4 sh = a.share[lo];
    if (sh.x == SHARE_ALIAS || (sh.x == SHARE_PREFIX && (ti + 1) * GT <= sh.z * BLK)) return;
  }
  const double* x = a.xg + m.xoff;
  const double* prm = a.prm + m.poff;
  double out[4][4];
  gram_tile(m.ktype, a.D, prm, x, m.np, ti * GT, m.n, x, m.np, tj * GT, m.n, out, sxa, sxb, scoef, sT);
  const double cnoise = prm[PRM_C];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double* F = a.F + m.foff;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = tj * GT + ty * 4 + j;
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = ti * GT + tx * 4 + i;
      double k = out[i][j];
      if (r >= m.n || c >= m.n) k = (r == c) ? 1.0 : 0.0;
      else if (r == c) k += cnoise;
      v[i] = k;
    }
    double* dst = F + tidx(ti * GT + tx * 4, c, m.nkc);
    *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
  }
}

// Rectangular Gram K(xa, xb) -> out (na x nb column-major, ld = ldo).  grid = (ceil(na/GT), ceil(nb/GT)).
__global__ void __launch_bounds__(NTHREADS) gram_rect_kernel(GramRectArgs a) {
  __shared__ double sxa[GDC * GT], sxb[GDC * GT], scoef[GDC], sT[EXPTAB_N];
  exptab_load(sT);
  double out[4][4];
  const int ra0 = blockIdx.x * GT, rb0 = blockIdx.y * GT;
  gram_tile(a.ktype, a.D, a.prm, a.xa, a.sa, ra0, a.na, a.xb, a.sb, rb0, a.nb, out, sxa, sxb, scoef, sT);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = rb0 + ty * 4 + j;
    if (c >= a.nb) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = ra0 + tx * 4 + i;
      if (r < a.na) a.out[(int64_t)c * a.ldo + r] = out[i][j];
    }
  }
}

// Gather the leaves' input rows into SoA blocks: xg[xoff + d*np + p] = x[(obs[p]-1) + d*N], zero padded.
__global__ void gather_kernel(GatherArgs a) {
  const LeafMeta m = a.meta[blockIdx.y];
  const int64_t* obs = a.obs + a.obs_off[blockIdx.y];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m.np; p += gridDim.x * blockDim.x) {
    for (int d = 0; d < a.D; d++)
      a.xg[m.xoff + (int64_t)d * m.np + p] = (p < m.n) ? a.x[(obs[p] - 1) + (int64_t)d * a.N] : 0.0;
  }
}

}  // namespace dsm
