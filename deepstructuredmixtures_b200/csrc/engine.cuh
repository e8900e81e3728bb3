// FP64 macro-tile contraction engine:  acc(128x128) += A(128xK) * B(128xK)^T  on DMMA.8x8x4.
//
// Every O(n^3) step of the hot path (left-looking Cholesky panels, TRSM-by-inverse, triangular
// inverse, LAUUM, predictive forward substitution) is this one contraction with different operand
// sources and epilogues.  Operands are column-major with the contraction index k as the column:
// element (r, k) lives at P[r + k*ld], so a k-slice of a tile is 128 contiguous doubles (coalesced,
// 16-byte cp.async).  smem tiles are [k][LDS] with LDS = 132 so the DMMA fragment loads
// (row = lane/4, k = lane%4) touch 16 distinct 8-byte bank pairs per half warp.
//
// 256 threads = 8 warps arranged 2 (M) x 4 (N); each warp owns a 64 x 32 sub tile =
// 8 x 4 DMMA tiles = 64 FP64 accumulators per thread.  Roofline: 37.2 TFLOP/s DMMA issue on B200
// (tools/fp64_peaks.cu), per CTA tile 16 flop per byte staged -> 2.3 TB/s of L2->SM traffic at peak.
#pragma once
#include "common.cuh"

namespace dsm {

typedef double Acc[8][4][2];

__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// RES = 0: A and B streamed from global (stages in `sA`/`sB`, stage strides strideA/strideB doubles)
// RES = 1: A resident in smem at sA as [k][LDS]; B streamed
// RES = 2: B resident in smem at sB as [k][LDS]; A streamed
// Operand element (r, k) of chunk c (k in [16c, 16c+16)) is at  P[c*cs + (k - 16c)*ld + r]: column-major operands
// use (ld, cs) = (lda, 16*lda), operands in the chunk-tiled layout (common.cuh) use (LDS, TILE_D).
// mrows / ncols: valid rows of A / B (64 or 128).  K: multiple of KC.
// tri = 1: skip warp tiles strictly above the diagonal (symmetric result, lower part wanted).
// klim_rows = 1 (RES==2 only): A is lower triangular in (r,k): rows of warp-half wm only need k < wm*64+64.
// klim_cols = 1 (RES==1 only): B is lower triangular in (c,k): cols of warp wn only need k < wn*32+32.
template <int RES>
__device__ __forceinline__ void mma_run(Acc& acc, const double* __restrict__ A, int64_t lda, int64_t csa,
                                        const double* __restrict__ B, int64_t ldb, int64_t csb, int K,
                                        int mrows, int ncols, bool tri,
                                        double* sA, int strideA, double* sB, int strideB,
                                        bool klim_rows = false, bool klim_cols = false) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const bool active = (wm * 64 < mrows) && (wn * 32 < ncols) && !(tri && wn * 32 >= wm * 64 + 64);
  int kmax = K;
  if (klim_rows) kmax = min(K, wm * 64 + 64);
  if (klim_cols) kmax = min(K, wn * 32 + 32);
  const int nch = K / KC;

  auto issue = [&](int c) {
    const int st = c % NST;
    if (RES != 1) {
      double* dst = sA + st * strideA;
      const double* src = A + (int64_t)c * csa;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int u = tid + q * NTHREADS, kk = u >> 6, r2 = (u & 63) << 1;
        if (r2 < mrows) cp_async16(dst + kk * LDS + r2, src + (int64_t)kk * lda + r2);
      }
    }
    if (RES != 2) {
      double* dst = sB + st * strideB;
      const double* src = B + (int64_t)c * csb;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int u = tid + q * NTHREADS, kk = u >> 6, r2 = (u & 63) << 1;
        if (r2 < ncols) cp_async16(dst + kk * LDS + r2, src + (int64_t)kk * ldb + r2);
      }
    }
  };

#pragma unroll
  for (int c = 0; c < NST - 1; c++) {
    if (c < nch) issue(c);
    cp_async_commit();
  }
  for (int c = 0; c < nch; c++) {
    cp_async_wait<NST - 2>();
    __syncthreads();
    if (c + NST - 1 < nch) issue(c + NST - 1);
    cp_async_commit();
    if (active && c * KC < kmax) {
      const double* pa = (RES == 1) ? sA + c * CHUNK : sA + (c % NST) * strideA;
      const double* pb = (RES == 2) ? sB + c * CHUNK : sB + (c % NST) * strideB;
      pa += (lane & 3) * LDS + wm * 64 + (lane >> 2);
      pb += (lane & 3) * LDS + wn * 32 + (lane >> 2);
#pragma unroll
      for (int ks = 0; ks < KC / 4; ks++) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = pa[ks * 4 * LDS + i * 8];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = pb[ks * 4 * LDS + j * 8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// Accumulator element (i, j, e) of this thread is tile element
//   row = wm*64 + i*8 + lane/4 ,  col = wn*32 + j*8 + 2*(lane%4) + e.
// Store as [col][LDS] (column-major tile: usable as a resident A operand with k = col).
__device__ __forceinline__ void acc_store_colmajor(const Acc& acc, double* S, double scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp & 1, wn = warp >> 1;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3);
      S[c * LDS + r] = scale * acc[i][j][0];
      S[(c + 1) * LDS + r] = scale * acc[i][j][1];
    }
}
// Store as [row][LDS] (row-major tile: usable as a resident B operand with k = row).
__device__ __forceinline__ void acc_store_rowmajor(const Acc& acc, double* S, double scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp & 1, wn = warp >> 1;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = wm * 64 + i * 8 + (lane >> 2), c = wn * 32 + j * 8 + 2 * (lane & 3);
      double2 v = make_double2(scale * acc[i][j][0], scale * acc[i][j][1]);
      *reinterpret_cast<double2*>(S + r * LDS + c) = v;
    }
}

// ---- in-shared-memory dense kernels on a resident tile S[c*LDS + r] (column-major, lower) ------

// Unblocked right-looking Cholesky of the leading m x m block.  Returns the LAPACK info of the block
// (0, or 1-based index of the first non-positive pivot; the factorisation then continues with NaNs,
// which is what the reference's unchecked potrf! + log() makes observable).  `sh` >= 2 doubles.
__device__ __forceinline__ int potrf_smem(double* S, int m, double* sh) {
  const int tid = threadIdx.x;
  int info = 0;
  for (int j = 0; j < m; j++) {
    if (tid == 0) {
      const double d = S[j * LDS + j];
      if (!(d > 0.0) && sh[1] == 0.0) sh[1] = (double)(j + 1);
      const double r = sqrt(d);
      S[j * LDS + j] = r;
      sh[0] = 1.0 / r;
    }
    __syncthreads();
    const double inv = sh[0];
    for (int r = j + 1 + tid; r < m; r += NTHREADS) S[j * LDS + r] *= inv;
    __syncthreads();
    // trailing update of columns j+1..m-1: thread -> (column c, row segment)
    const int rem = m - j - 1;
    if (rem > 0) {
      // 256 threads: 2 threads per column when rem <= 128 (always true)
      const int c = j + 1 + (tid >> 1);
      if (c < m) {
        const double ljc = S[j * LDS + c];
        const double* lj = S + j * LDS;
        double* sc = S + c * LDS;
        for (int r = c + (tid & 1); r < m; r += 2) sc[r] -= lj[r] * ljc;
      }
    }
    __syncthreads();
  }
  info = (int)sh[1];
  return info;
}

// In-place inverse of the lower-triangular m x m block (LAPACK dtrti2 'L','N' order, backwards):
// W[j+1:,j] = -W[j+1:,j+1:] * L[j+1:,j] / L[j,j].  `col` >= BLK doubles of smem.
__device__ __forceinline__ void trtri_smem(double* S, int m, double* col) {
  const int tid = threadIdx.x;
  for (int j = m - 1; j >= 0; j--) {
    for (int r = j + tid; r < m; r += NTHREADS) col[r] = S[j * LDS + r];
    __syncthreads();
    const double ajj = 1.0 / col[j];
    // row r (j < r < m):  sum_{k=j+1..r} W[r][k] * L[k][j]   (2 threads per row, even/odd k)
    const int r = j + 1 + (tid >> 1);
    double s = 0.0;
    if (r < m) {
      for (int k = j + 1 + (tid & 1); k <= r; k += 2) s += S[k * LDS + r] * col[k];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (r < m && (tid & 1) == 0) S[j * LDS + r] = -s * ajj;
    if (tid == 0) S[j * LDS + j] = ajj;
    __syncthreads();
  }
}

}  // namespace dsm
