// FP64 macro-tile contraction engine v2:  OUT(128 x 128) += A(128 x K) * B(128 x K)^T  on DMMA.8x8x4.
//
// Warp w owns the ROW SLAB  rows [16w, 16w+16) x all 128 columns  of the macro tile (2 x 16 DMMA tiles = 64 FP64
// accumulators per thread).  Because a warp holds complete rows, the triangular right-multiplication that follows
// every contraction in this code base (TRSM by the inverse diagonal block:  OUT <- OUT * M^T,  M lower triangular)
// runs entirely in registers: OUT's accumulator fragments are re-shaped into A fragments with quad shuffles and the
// only shared-memory operand is M, streamed through the same ring as everything else.  No staging tile, no block
// barrier.  (v1 staged OUT through a 135 KB shared tile and synchronised the block twice per tile.)
//
// Accumulator element (mb, nb, e) of thread (g = lane/4, t = lane%4) in warp w -- INTERLEAVED mapping:
//     row = 16 slab(w) + 2 g + mb                (the two m-tiles own the even / odd rows of the slab)
//     col = 16 (nb/2) + 2 (2 t + e) + (nb % 2)   (tiles 2m, 2m+1 own the even / odd columns of a 16-column group)
// so that the fragments of a tile PAIR are adjacent in shared memory and one LDS.128 feeds two DMMAs' operands:
// 9 shared loads per k-step (1 for A, 8 for B) instead of 18.  Any consistent permutation of rows / columns is
// legal for a contraction; only the stores and the epilogues have to agree with it.
#pragma once
#include "pipe.cuh"

namespace dsm {

typedef double Acc2[2][16][2];

// Row slab of an MMA warp.  Warps w and w + 4 share an SM sub-partition (and its DMMA pipe); slabs are dealt so that the
// two warps of a sub-partition own slabs s and 7 - s: whenever work is triangular in the slab index (the lower triangle
// of a diagonal tile, the structurally-zero part of a triangular operand) every sub-partition gets the same share.
//   warp 0 1 2 3 4 5 6 7  ->  slab 0 2 4 6 7 5 3 1
__device__ __forceinline__ int warp_slab() { const int w = threadIdx.x >> 5; return w < 4 ? 2 * w : 15 - 2 * w; }
__device__ __forceinline__ int acc_row(int mb) { return 16 * warp_slab() + 2 * ((threadIdx.x & 31) >> 2) + mb; }
__device__ __forceinline__ int acc_col(int nb, int e) { return 16 * (nb >> 1) + 2 * (2 * (threadIdx.x & 3) + e) + (nb & 1); }

__device__ __forceinline__ void acc2_zero(Acc2& acc) {
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int n = 0; n < 16; n++) { acc[m][n][0] = 0.0; acc[m][n][1] = 0.0; }
}

// One 16-k chunk of the contraction; N16 = number of live 16-column groups (cols < 16*N16).
template <int N16>
__device__ __forceinline__ void mma_chunk16(Acc2& acc, const double* __restrict__ sA, const double* __restrict__ sB, int r0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double* pa = sA + t * LDS + r0 + 2 * g;
  const double* pb = sB + t * LDS + 2 * g;
#pragma unroll
  for (int ks = 0; ks < KC / 4; ks++) {
    const double2 a = *reinterpret_cast<const double2*>(pa + ks * 4 * LDS);
#pragma unroll
    for (int m2 = 0; m2 < N16; m2++) {
      const double2 b = *reinterpret_cast<const double2*>(pb + ks * 4 * LDS + 16 * m2);
      dmma884(acc[0][2 * m2][0], acc[0][2 * m2][1], a.x, b.x);
      dmma884(acc[1][2 * m2][0], acc[1][2 * m2][1], a.y, b.x);
      dmma884(acc[0][2 * m2 + 1][0], acc[0][2 * m2 + 1][1], a.x, b.y);
      dmma884(acc[1][2 * m2 + 1][0], acc[1][2 * m2 + 1][1], a.y, b.y);
    }
  }
}

// One 16-k chunk of the contraction; NG = number of 4-tile column groups that are live (cols < 32*NG).
template <int NG>
__device__ __forceinline__ void mma_chunk(Acc2& acc, const double* __restrict__ sA, const double* __restrict__ sB, int r0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double* pa = sA + t * LDS + r0 + 2 * g;
  const double* pb = sB + t * LDS + 2 * g;
#pragma unroll
  for (int ks = 0; ks < KC / 4; ks++) {
    const double2 a = *reinterpret_cast<const double2*>(pa + ks * 4 * LDS);
#pragma unroll
    for (int m2 = 0; m2 < 2 * NG; m2++) {
      const double2 b = *reinterpret_cast<const double2*>(pb + ks * 4 * LDS + 16 * m2);
      dmma884(acc[0][2 * m2][0], acc[0][2 * m2][1], a.x, b.x);
      dmma884(acc[1][2 * m2][0], acc[1][2 * m2][1], a.y, b.x);
      dmma884(acc[0][2 * m2 + 1][0], acc[0][2 * m2 + 1][1], a.x, b.y);
      dmma884(acc[1][2 * m2 + 1][0], acc[1][2 * m2 + 1][1], a.y, b.y);
    }
  }
}

// acc -= T for the 16-column group m2 of the macro tile, T = a [16][LDS] tile image in shared memory (columns
// 16 m2 .. 16 m2 + 15 of the 128 x 128 block, rows = block rows).
__device__ __forceinline__ void acc2_sub_tile(Acc2& acc, const double* __restrict__ sT, int r0, const int M2) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int p = 0; p < 2; p++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const double2 v = *reinterpret_cast<const double2*>(sT + (2 * (2 * t + e) + p) * LDS + r0 + 2 * g);
      acc[0][2 * M2 + p][e] -= v.x;
      acc[1][2 * M2 + p][e] -= v.y;
    }
}

// A-fragment (row = lane/4, k = lane%4) for k = 16 J + 4 ks + t, taken from the accumulator tiles 2J / 2J+1 of m-tile mb:
// column k lives in tile 2J + (t & 1), lane t'' = ks, element e = t >> 1.
template <int J>
__device__ __forceinline__ double acc_to_afrag(const Acc2& acc, int mb, int ks) {
  const int lane = threadIdx.x & 31;
  const int src = (lane & ~3) | ks;
  const double v00 = __shfl_sync(0xffffffffu, acc[mb][2 * J][0], src);
  const double v01 = __shfl_sync(0xffffffffu, acc[mb][2 * J][1], src);
  const double v10 = __shfl_sync(0xffffffffu, acc[mb][2 * J + 1][0], src);
  const double v11 = __shfl_sync(0xffffffffu, acc[mb][2 * J + 1][1], src);
  const double ve = (lane & 2) ? v01 : v00, vo = (lane & 2) ? v11 : v10;
  return (lane & 1) ? vo : ve;
}

// One epilogue k-tile: pass P (output columns [32P, 32P+32)), k-tile J (k in [16J, 16J+16)):  sM[k][c] = M[c][16J + k].
// (the last k-tile of a pass, J = 2P+1, holds k in [32P+16, 32P+32): M[c][k] = 0 for the first 16 output columns)
template <int P, int J>
__device__ __forceinline__ void tri_chunk(const Acc2& acc, double (&o)[2][4][2], const double* __restrict__ sM) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double* pb = sM + t * LDS + 32 * P + 2 * g;
#pragma unroll
  for (int ks = 0; ks < 4; ks++) {
    const double a0 = acc_to_afrag<J>(acc, 0, ks);
    const double a1 = acc_to_afrag<J>(acc, 1, ks);
#pragma unroll
    for (int m2 = (J == 2 * P + 1) ? 1 : 0; m2 < 2; m2++) {
      const double2 b = *reinterpret_cast<const double2*>(pb + ks * 4 * LDS + 16 * m2);
      dmma884(o[0][2 * m2][0], o[0][2 * m2][1], a0, b.x);
      dmma884(o[1][2 * m2][0], o[1][2 * m2][1], a1, b.x);
      dmma884(o[0][2 * m2 + 1][0], o[0][2 * m2 + 1][1], a0, b.y);
      dmma884(o[1][2 * m2 + 1][0], o[1][2 * m2 + 1][1], a1, b.y);
    }
  }
}

// Triangular epilogue  OUT <- scale * OUT * M^T  (M lower triangular, 32*NP x 32*NP) in registers.
// M arrives as NP epilogue stages in DESCENDING k order: stage e carries M's k-tiles 2(NP-1-e) (A part) and
// 2(NP-1-e)+1 (B part).  The passes run over descending column groups (so the update is in place) and pass P reads
// k-tiles 0 .. 2P+1 only: the stage with the highest k-tiles is dead after the first pass, the next one after the
// second, ... -- they are released in ring (FIFO) order as the passes retire, so the producer can refill the ring with
// the next task's chunks while the remaining passes still run.
template <int P, int J>
struct TriPass {
  __device__ static __forceinline__ void run(const Pipe& p, uint32_t qlast, const Acc2& acc, double (&o)[2][4][2]) {
    TriPass<P, J - 1>::run(p, qlast, acc, o);
    const int st = (qlast - (J >> 1)) % NS2;
    tri_chunk<P, J>(acc, o, (J & 1) ? p.B(st) : p.A(st));
  }
};
template <int P>
struct TriPass<P, -1> {
  __device__ static __forceinline__ void run(const Pipe&, uint32_t, const Acc2&, double (&)[2][4][2]) {}
};

template <int P>
__device__ __forceinline__ void tri_pass(const Pipe& p, uint32_t qlast, Acc2& acc, double scale) {
  double o[2][4][2];
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int n = 0; n < 4; n++) { o[m][n][0] = 0.0; o[m][n][1] = 0.0; }
  TriPass<P, 2 * P + 1>::run(p, qlast, acc, o);
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int n = 0; n < 4; n++) { acc[m][4 * P + n][0] = scale * o[m][n][0]; acc[m][4 * P + n][1] = scale * o[m][n][1]; }
}

__host__ __device__ __forceinline__ int tri_epilogue_nstages(int npass) { return npass; }   // 8 (4) k-tiles, 2 per stage

__device__ __forceinline__ void tri_epilogue(Pipe& p, Acc2& acc, int npass, bool active, double scale) {
  const uint32_t q0 = p.q_cons;
  const uint32_t qlast = q0 + npass - 1;            // chunk that carries k-tiles 0 and 1
  for (int e = 0; e < npass; e++) {                 // pass NP-1 needs every stage
    const uint32_t q = q0 + e;
    p.wait_bar(&p.full[q % NS2], (q / NS2) & 1, 4);
  }
#if DSM_EPI_EARLY
  if (npass == 4) {
    if (active) tri_pass<3>(p, qlast, acc, scale);
    p.release();
    if (active) tri_pass<2>(p, qlast, acc, scale);
    p.release();
  }
  if (active) tri_pass<1>(p, qlast, acc, scale);
  p.release();
  if (active) tri_pass<0>(p, qlast, acc, scale);
  p.release();
#else
  if (active) {
    if (npass == 4) { tri_pass<3>(p, qlast, acc, scale); tri_pass<2>(p, qlast, acc, scale); }
    tri_pass<1>(p, qlast, acc, scale);
    tri_pass<0>(p, qlast, acc, scale);
  }
  for (int e = 0; e < npass; e++) p.release();
#endif
}

// Descriptor of epilogue stage `e` of `npass` for a tiled W block (k-tiles 2(npass-1-e) and 2(npass-1-e)+1).
__device__ __forceinline__ ChunkDesc tri_epilogue_chunk(const double* Wblk, int e, int npass, const int* flag) {
  ChunkDesc d;
  const int kt = 2 * (npass - 1 - e);
  d.a = Wblk + kt * TILE_D; d.abytes = TILE_BYTES;
  d.b = Wblk + (kt + 1) * TILE_D; d.bbytes = TILE_BYTES;
  d.flag0 = (e == 0) ? flag : nullptr; d.flag1 = nullptr;
  return d;
}

// Store OUT (rows r < mrows, cols c < ncols) into the tiled matrix at (row0 + r, col0 + c).
__device__ __forceinline__ void acc2_store(const Acc2& acc, double* Fm, int nkc, int row0, int col0, int mrows, int ncols) {
  const int r0 = 16 * warp_slab();
  if (r0 >= mrows) return;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    if (8 * n < ncols) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double* dst = Fm + tidx(row0 + acc_row(0), col0 + acc_col(n, e), nkc);        // rows 2g, 2g+1 are adjacent
        *reinterpret_cast<double2*>(dst) = make_double2(acc[0][n][e], acc[1][n][e]);
      }
    }
  }
}

}  // namespace dsm
