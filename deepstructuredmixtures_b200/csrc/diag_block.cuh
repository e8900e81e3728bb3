// Cholesky factor and inverse of ONE diagonal macro block (m x m, m = 64 or 128) resident in shared memory.
//
// This is the serial link of the left-looking factorisation (every panel tile of a block column waits for it), so
// it is blocked for latency, not for throughput:
//   * 32-wide panels.  The 32 x 32 diagonal block of a panel is factored AND inverted by one warp entirely in
//     registers (lane r holds row r; the rank-1 updates fetch the pivot column with warp shuffles) -- no shared-memory
//     round trips and no block barriers inside the 32 pivot steps.
//   * panel TRSM (A21 * D^-T), trailing SYRK and the block forward substitution that builds W = L^-1 run on
//     DMMA.8x8x4 with operands straight from the tile.
// Layout: S[c * LDS + r] column-major.  On return the lower triangle (incl. diagonal) holds L, the strict upper
// triangle of the OFF-diagonal 32-blocks holds W^T (W[r][c] at S[r * LDS + c]) and DI[P] holds the inverse of the
// P-th 32 x 32 diagonal block as DI[P][k * DLD + x] = W[x][k] (zero for x < k).
#pragma once
#include "common.cuh"

namespace dsm {

constexpr int PW = 32;            // panel width
constexpr int DLD = 36;           // leading dimension of the 32 x 32 scratch blocks (36 % 16 == 4: conflict-free fragments)
constexpr int DBLK = PW * DLD;    // doubles per scratch block
constexpr int DIAG_AUX_DOUBLES = 4 * DBLK + 3 * DBLK;   // DI[4] + T[3]

// W[r][k] (r >= k) after diag_factor_invert
__device__ __forceinline__ double diag_W(const double* S, const double* DI, int r, int k) {
  if ((r / PW) == (k / PW)) return DI[(r / PW) * DBLK + (k % PW) * DLD + (r % PW)];
  return S[r * LDS + k];
}

// Pivot step j of the register-resident 32 x 32 Cholesky (lane r holds row r in a[]).  Template recursion forces the
// complete unrolling that keeps a[] in registers (register arrays cannot be indexed dynamically).
template <int J>
struct PotrfStep {
  __device__ static __forceinline__ void run(double (&a)[PW], int r, int& info) {
    PotrfStep<J - 1>::run(a, r, info);
    const double d = __shfl_sync(0xffffffffu, a[J], J);
    if (!(d > 0.0) && info == 0) info = J + 1;
    const double inv = rsqrt(d);                 // one long-latency op on the pivot chain instead of sqrt + divide
    const double piv = d * inv;
    a[J] = (r == J) ? piv : a[J] * inv;
#pragma unroll
    for (int c = J + 1; c < PW; c++) {
      const double lcj = __shfl_sync(0xffffffffu, a[J], c);
      a[c] = fma(-a[J], lcj, a[c]);
    }
  }
};
template <>
struct PotrfStep<-1> {
  __device__ static __forceinline__ void run(double (&)[PW], int, int&) {}
};

// Row q of W = L^-1 for every column at once (lane r builds column r: w[k] = W[k][r], zero for k < r).
template <int Q>
struct InvStep {
  __device__ static __forceinline__ void run(const double (&a)[PW], double (&w)[PW], int r, double rinv) {
    InvStep<Q - 1>::run(a, w, r, rinv);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < Q; k++) {
      const double lqk = __shfl_sync(0xffffffffu, a[k], Q);
      if (k & 1) s1 = fma(lqk, w[k], s1); else s0 = fma(lqk, w[k], s0);
    }
    const double rq = __shfl_sync(0xffffffffu, rinv, Q);
    w[Q] = (r == Q) ? rq : ((r < Q) ? -(s0 + s1) * rq : 0.0);
  }
};
template <>
struct InvStep<-1> {
  __device__ static __forceinline__ void run(const double (&)[PW], double (&)[PW], int, double) {}
};

// Warp-level Cholesky + inverse of the 32 x 32 block at (c0, c0).  Executed by ONE warp.  Returns the 1-based local
// pivot index of the first non-positive pivot (0 = ok).
__device__ __forceinline__ int warp_potrf_inv32(double* S, int c0, double* DIp, bool factor) {
  const int r = threadIdx.x & 31;
  double a[PW];
#pragma unroll
  for (int c = 0; c < PW; c++) a[c] = S[(c0 + c) * LDS + c0 + r];      // row r (entries c > r are never used)
  int info = 0;
  if (factor) {
    PotrfStep<PW - 1>::run(a, r, info);
#pragma unroll
    for (int c = 0; c < PW; c++)
      if (r >= c) S[(c0 + c) * LDS + c0 + r] = a[c];
  }
  double w[PW];
  double rinv = 1.0;
#pragma unroll
  for (int k = 0; k < PW; k++) { w[k] = 0.0; if (k == r) rinv = 1.0 / a[k]; }     // 1 / l_rr without dynamic indexing
  InvStep<PW - 1>::run(a, w, r, rinv);
#pragma unroll
  for (int k = 0; k < PW; k++) DIp[r * DLD + k] = w[k];   // DI[k = column r][x = k] = W[k][r]  -> row `r` of the scratch
  return info;
}

__device__ __forceinline__ void dmma_tile(double& c0, double& c1, double a, double b) { dmma884(c0, c1, a, b); }

// S (m x m) -> L, W as described above.  All NTHREADS threads.  `aux` >= DIAG_AUX_DOUBLES doubles of shared memory.
// Returns the 1-based pivot index (within the block) of the first non-positive pivot, or 0.
__device__ __forceinline__ int diag_factor_invert(double* S, int m, double* aux, bool factor, int* s_info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  double* DI = aux;
  double* T = aux + 4 * DBLK;
  const int np = m / PW;
  if (tid == 0) *s_info = 0;
  __syncthreads();
  for (int P = 0; P < np; P++) {
    const int c0 = P * PW, r1 = c0 + PW, nrem = m - r1;
    if (warp == 0) {
      const int info = warp_potrf_inv32(S, c0, DI + P * DBLK, factor);
      if (lane == 0 && info != 0 && *s_info == 0) *s_info = c0 + info;
    }
    __syncthreads();
    if (!factor || nrem == 0) continue;
    // ---- panel TRSM: X = A21 * DI_P^T  (rows r1..m, 32 columns), in place
    for (int rt = warp; rt < nrem / 8; rt += NTHREADS / 32) {
      const int rr = r1 + 8 * rt;
      double a[8];
#pragma unroll
      for (int kb = 0; kb < 8; kb++) a[kb] = S[(c0 + 4 * kb + t) * LDS + rr + g];
      double x[4][2];
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        x[nb][0] = 0.0; x[nb][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < 8; kb++)
          if (kb <= 2 * nb + 1) dmma_tile(x[nb][0], x[nb][1], a[kb], DI[P * DBLK + (4 * kb + t) * DLD + 8 * nb + g]);
      }
      __syncwarp();
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        S[(c0 + 8 * nb + 2 * t) * LDS + rr + g] = x[nb][0];
        S[(c0 + 8 * nb + 2 * t + 1) * LDS + rr + g] = x[nb][1];
      }
    }
    __syncthreads();
    // ---- trailing update: A22 -= X X^T  (lower 8 x 8 tiles)
    {
      const int nt = nrem / 8, ntiles = nt * (nt + 1) / 2;
      for (int q = warp; q < ntiles; q += NTHREADS / 32) {
        int ti = (int)((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= q) ti++;
        while (ti * (ti + 1) / 2 > q) ti--;
        const int tj = q - ti * (ti + 1) / 2;
        double* cp = S + (r1 + 8 * tj + 2 * t) * LDS + r1 + 8 * ti + g;
        double c0v = cp[0], c1v = cp[LDS];
#pragma unroll
        for (int kb = 0; kb < 8; kb++) {
          const double av = -S[(c0 + 4 * kb + t) * LDS + r1 + 8 * ti + g];
          const double bv = S[(c0 + 4 * kb + t) * LDS + r1 + 8 * tj + g];
          dmma_tile(c0v, c1v, av, bv);
        }
        cp[0] = c0v; cp[LDS] = c1v;
      }
    }
    __syncthreads();
  }
  // ---- W = L^-1 by block forward substitution: row block P, column blocks j < P
  for (int P = 1; P < np; P++) {
    // T_j = sum_{k = 32 j}^{32 P - 1} L[32P + r][k] * W[k][32 j + c]      (16 tiles per j)
    for (int q = warp; q < 16 * P; q += NTHREADS / 32) {
      const int j = q >> 4, mb = (q >> 2) & 3, nb = q & 3;
      double c0v = 0.0, c1v = 0.0;
      for (int kb = 0; kb < 8 * (P - j); kb++) {
        const int kg = PW * j + 4 * kb + t;                       // global k of this lane's fragment element
        const double av = S[kg * LDS + PW * P + 8 * mb + g];
        const double bv = (kb < 8) ? DI[j * DBLK + (8 * nb + g) * DLD + 4 * kb + t]     // W[k][c] inside diagonal block j
                                   : S[kg * LDS + PW * j + 8 * nb + g];                  // W^T stored in the upper part
        dmma_tile(c0v, c1v, av, bv);
      }
      double* tp = T + j * DBLK + (8 * mb + g) * DLD + 8 * nb + 2 * t;   // T[k = row][c]
      tp[0] = c0v; tp[1] = c1v;
    }
    __syncthreads();
    // W_Pj = -DI_P * T_j   -> stored transposed: W[r][c] at S[r_global * LDS + c_global]
    for (int q = warp; q < 16 * P; q += NTHREADS / 32) {
      const int j = q >> 4, mb = (q >> 2) & 3, nb = q & 3;
      double c0v = 0.0, c1v = 0.0;
#pragma unroll
      for (int kb = 0; kb < 8; kb++)
        if (kb <= 2 * mb + 1)
          dmma_tile(c0v, c1v, DI[P * DBLK + (4 * kb + t) * DLD + 8 * mb + g], T[j * DBLK + (4 * kb + t) * DLD + 8 * nb + g]);
      double* wp = S + (PW * P + 8 * mb + g) * LDS + PW * j + 8 * nb + 2 * t;
      wp[0] = -c0v; wp[1] = -c1v;
    }
    __syncthreads();
  }
  return *s_info;
}

}  // namespace dsm
