// Cholesky factor and inverse of ONE diagonal macro block (m x m, m = 64 or 128) resident in shared memory.
//
// This is the serial link of the left-looking factorisation (every panel tile of a block column waits for it), so
// it is blocked for latency, not for throughput:
//   * 16-wide panels.  The 16 x 16 diagonal block of a panel is factored AND inverted by one warp entirely in
//     registers (lane r holds row r; the rank-1 updates fetch the pivot column with warp shuffles) -- no shared-memory
//     round trips and no block barriers inside the pivot steps.  (The register routine costs O(PW^2) shuffles per
//     panel, i.e. O(m PW) per tile: 16 beat 32 by 3x in the per-task clock stamps.)
//   * panel TRSM (A21 * D^-T) and the trailing SYRK run on DMMA.8x8x4 with operands straight from the tile.
//   * W = L^-1 by block forward substitution, one WARP PER BLOCK COLUMN: column j of W depends only on L and on
//     itself, so the eight columns proceed without any block barrier (DMMA + a per-warp 16 x 16 scratch).
// Layout: S[c * LDS + r] column-major.  On return the lower triangle (incl. diagonal) holds L, the strict upper
// triangle of the OFF-diagonal 16-blocks holds W^T (W[r][c] at S[r * LDS + c]) and DI[P] holds the inverse of the
// P-th 16 x 16 diagonal block as DI[P][k * DLD + x] = W[x][k] (zero for x < k).
#pragma once
#include "common.cuh"

namespace dsm {

constexpr int PW = 16;            // panel width
constexpr int DLD = 20;           // leading dimension of the 16 x 16 scratch blocks (20 % 16 == 4: conflict-free fragments)
constexpr int DBLK = PW * DLD;    // doubles per scratch block
constexpr int DIAG_AUX_DOUBLES = (BLK / PW) * DBLK + (NTHREADS / 32) * DBLK + BLK;   // DI[8] + one T block per warp + 1/diag

// W[r][k] (r >= k) after diag_factor_invert
__device__ __forceinline__ double diag_W(const double* S, const double* DI, int r, int k) {
  if ((r / PW) == (k / PW)) return DI[(r / PW) * DBLK + (k % PW) * DLD + (r % PW)];
  return S[r * LDS + k];
}

// Pivot step j of the register-resident 32 x 32 Cholesky (lane r holds row r in a[]).  Template recursion forces the
// complete unrolling that keeps a[] in registers (register arrays cannot be indexed dynamically).
template <int J>
struct PotrfStep {
  __device__ static __forceinline__ void run(double (&a)[PW], int r, int& info, double& myinv) {
    PotrfStep<J - 1>::run(a, r, info, myinv);
    const double d = __shfl_sync(0xffffffffu, a[J], J);
    if (!(d > 0.0) && info == 0) info = J + 1;
    const double inv = rsqrt(d);                 // one long-latency op on the pivot chain instead of sqrt + divide
    const double piv = d * inv;
    if (r == J) myinv = inv;                     // 1 / l_JJ for the panel TRSM and the block inverse: no division anywhere
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
  __device__ static __forceinline__ void run(double (&)[PW], int, int&, double&) {}
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

// Warp-level Cholesky of the PW x PW block at (c0, c0) (lanes 16..31 mirror lanes 0..15): L is written back to S,
// the reciprocals of its diagonal to rdiag[c0 ..].  Returns the 1-based local index of the first non-positive pivot.
__device__ __forceinline__ int warp_potrf16(double* S, int c0, double* rdiag) {
  const int r = threadIdx.x & (PW - 1);
  const bool writer = (threadIdx.x & 31) < PW;
  double a[PW];
#pragma unroll
  for (int c = 0; c < PW; c++) a[c] = S[(c0 + c) * LDS + c0 + r];      // row r (entries c > r are never used)
  int info = 0;
  double myinv = 1.0;
  PotrfStep<PW - 1>::run(a, r, info, myinv);
#pragma unroll
  for (int c = 0; c < PW; c++)
    if (writer && r >= c) S[(c0 + c) * LDS + c0 + r] = a[c];
  if (writer) rdiag[c0 + r] = myinv;
  return info;
}

// Warp-level inverse of the lower-triangular PW x PW block at (c0, c0) into the scratch block DIp.
// rdiag: reciprocals of the diagonal from the factorisation, or null (a block that was not factored here).
__device__ __forceinline__ void warp_inv16(const double* S, int c0, double* DIp, const double* rdiag) {
  const int r = threadIdx.x & (PW - 1);
  const bool writer = (threadIdx.x & 31) < PW;
  double a[PW];
#pragma unroll
  for (int c = 0; c < PW; c++) a[c] = S[(c0 + c) * LDS + c0 + r];
  double w[PW];
  double rinv = 1.0;
#pragma unroll
  for (int k = 0; k < PW; k++) w[k] = 0.0;
  if (rdiag != nullptr) rinv = rdiag[c0 + r];
  else {
#pragma unroll
    for (int k = 0; k < PW; k++) if (k == r) rinv = 1.0 / a[k];                    // 1 / l_rr without dynamic indexing
  }
  InvStep<PW - 1>::run(a, w, r, rinv);
#pragma unroll
  for (int k = 0; k < PW; k++)
    if (writer) DIp[r * DLD + k] = w[k];                  // DI[k = column r][x = k] = W[k][r]  -> row `r` of the scratch
}

__device__ __forceinline__ void dmma_tile(double& c0, double& c1, double a, double b) { dmma884(c0, c1, a, b); }

// S (m x m) -> L, W as described above.  All NTHREADS threads.  `aux` >= DIAG_AUX_DOUBLES doubles of shared memory.
// Returns the 1-based pivot index (within the block) of the first non-positive pivot, or 0.
__device__ __forceinline__ int diag_factor_invert(double* S, int m, double* aux, bool factor, int* s_info, long long* tstamp = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  double* DI = aux;
  double* Tw = aux + (BLK / PW) * DBLK + warp * DBLK;       // this warp's 16 x 16 scratch
  const int np = m / PW;
  if (tid == 0) *s_info = 0;
  csync();
  double* rdiag = aux + (BLK / PW) * DBLK + (NTHREADS / 32) * DBLK;     // 1 / l_cc of the tile
  // Look-ahead: the 16 x 16 pivot block of panel P+1 is updated first and factored by warp 0 WHILE the other warps
  // finish the trailing update of panel P, so the serial register Cholesky leaves the critical path of the tile.
  if (factor && warp == 0) {
    const int info = warp_potrf16(S, 0, rdiag);
    if (lane == 0 && info != 0 && *s_info == 0) *s_info = info;
  }
  csync();
  for (int P = 0; P < np && factor; P++) {
    const int c0 = P * PW, r1 = c0 + PW, nrem = m - r1;
    if (nrem == 0) continue;
    // ---- panel TRSM: X = A21 * L11^-T by forward substitution, one thread per row (needs L11 only, not its inverse)
    if (tid < nrem) {
      const int rr = r1 + tid;
      double x[PW];
#pragma unroll
      for (int c = 0; c < PW; c++) x[c] = S[(c0 + c) * LDS + rr];
#pragma unroll
      for (int c = 0; c < PW; c++) {
        x[c] *= rdiag[c0 + c];
#pragma unroll
        for (int k = c + 1; k < PW; k++) x[k] = fma(-x[c], S[(c0 + c) * LDS + c0 + k], x[k]);     // l_kc (broadcast read)
      }
#pragma unroll
      for (int c = 0; c < PW; c++) S[(c0 + c) * LDS + rr] = x[c];
    }
    csync();
    // ---- trailing update: A22 -= X X^T  (lower 8 x 8 tiles).  Tiles 0..2 are the next pivot block: warp 0 takes them,
    // then factors that block; warps 1..7 share the rest.
    {
      const int nt = nrem / 8, ntiles = nt * (nt + 1) / 2;
      const int qbeg = (warp == 0) ? 0 : 3 + (warp - 1), qend = (warp == 0) ? min(3, ntiles) : ntiles;
      const int qstep = (warp == 0) ? 1 : (NCONS / 32 - 1);
      for (int q = qbeg; q < qend; q += qstep) {
        int ti = (int)((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= q) ti++;
        while (ti * (ti + 1) / 2 > q) ti--;
        const int tj = q - ti * (ti + 1) / 2;
        double* cp = S + (r1 + 8 * tj + 2 * t) * LDS + r1 + 8 * ti + g;
        double c0v = cp[0], c1v = cp[LDS];
#pragma unroll
        for (int kb = 0; kb < PW / 4; kb++) {
          const double av = -S[(c0 + 4 * kb + t) * LDS + r1 + 8 * ti + g];
          const double bv = S[(c0 + 4 * kb + t) * LDS + r1 + 8 * tj + g];
          dmma_tile(c0v, c1v, av, bv);
        }
        cp[0] = c0v; cp[LDS] = c1v;
      }
      if (warp == 0) {
        __syncwarp();
        const int info = warp_potrf16(S, r1, rdiag);
        if (lane == 0 && info != 0 && *s_info == 0) *s_info = r1 + info;
      }
    }
    csync();
  }
  if (tstamp) tstamp[0] = clock64();
  // ---- inverses of the diagonal blocks: eight independent 16 x 16 problems, one warp each
  for (int P = warp; P < np; P += NTHREADS / 32) warp_inv16(S, P * PW, DI + P * DBLK, factor ? rdiag : nullptr);
  csync();
  if (tstamp) tstamp[1] = clock64();
  // ---- W = L^-1: warp j owns block column j; row blocks P = j+1 .. np-1 in sequence, no block barrier
  for (int j = warp; j < np; j += NTHREADS / 32) {
    for (int P = j + 1; P < np; P++) {
      // T = sum_{k = PW j}^{PW P - 1} L[PW P + r][k] * W[k][PW j + c]        (2 x 2 tiles)
      double c[2][2][2];
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
#pragma unroll
        for (int nb = 0; nb < 2; nb++) { c[mb][nb][0] = 0.0; c[mb][nb][1] = 0.0; }
      for (int kb = 0; kb < (PW / 4) * (P - j); kb++) {
        const int kg = PW * j + 4 * kb + t;
        const double a0 = S[kg * LDS + PW * P + g], a1 = S[kg * LDS + PW * P + 8 + g];
        double b0, b1;
        if (kb < PW / 4) {            // W[k][c] inside diagonal block j
          b0 = DI[j * DBLK + g * DLD + 4 * kb + t];
          b1 = DI[j * DBLK + (8 + g) * DLD + 4 * kb + t];
        } else {                       // W^T stored in the upper part (written by this warp in earlier iterations)
          b0 = S[kg * LDS + PW * j + g];
          b1 = S[kg * LDS + PW * j + 8 + g];
        }
        dmma_tile(c[0][0][0], c[0][0][1], a0, b0);
        dmma_tile(c[0][1][0], c[0][1][1], a0, b1);
        dmma_tile(c[1][0][0], c[1][0][1], a1, b0);
        dmma_tile(c[1][1][0], c[1][1][1], a1, b1);
      }
      __syncwarp();
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
#pragma unroll
        for (int nb = 0; nb < 2; nb++) {
          double* tp = Tw + (8 * mb + g) * DLD + 8 * nb + 2 * t;         // T[k = row][c]
          tp[0] = c[mb][nb][0]; tp[1] = c[mb][nb][1];
        }
      __syncwarp();
      // W_Pj = -DI_P * T  -> stored transposed: W[r][c] at S[r_global * LDS + c_global]
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
#pragma unroll
        for (int nb = 0; nb < 2; nb++) {
          double w0 = 0.0, w1 = 0.0;
#pragma unroll
          for (int kb = 0; kb < PW / 4; kb++)
            if (kb <= 2 * mb + 1)
              dmma_tile(w0, w1, DI[P * DBLK + (4 * kb + t) * DLD + 8 * mb + g], Tw[(4 * kb + t) * DLD + 8 * nb + g]);
          double* wp = S + (PW * P + 8 * mb + g) * LDS + PW * j + 8 * nb + 2 * t;
          wp[0] = -w0; wp[1] = -w1;
        }
      __syncwarp();
    }
  }
  csync();
  return *s_info;
}

}  // namespace dsm
