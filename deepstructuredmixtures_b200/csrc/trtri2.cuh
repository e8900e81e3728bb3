// Batched triangular inverse X = L^-1 (engine v2) with fused  tr(F^-1) = ||X||_F^2  and  alpha = X^T z.
//
// Replaces the n-RHS ldiv!(cK, -I) of ααinvcK! (gaussianprocess.jl:219-226) and the backward solve of
// gaussianprocess.jl:105.  Task = (leaf, block column J), no cross-task dependency: for I = J+1 .. nb-1
//     S   = sum_{K=J}^{I-1} L_IK X_KJ            (K = J uses X_JJ = W_J)
//     X_IJ = -W_I S
// The engine computes the TRANSPOSE  OUT[c][r] = S^T  (A operand = X^T rows of block J, B operand = L rows of block I)
// so that a warp owns complete rows c and the right-multiplication by W_I^T runs in registers; X_IJ^T is stored in
// the strict upper block triangle of the factor (rows of block J, columns of block I), which is also where the next
// iteration's A operand is read from.  The producer warp therefore may run ahead of the consumers only up to the
// k-block written by the previous iteration (`stored`).
#pragma once
#include "engine2.cuh"
#include "args.h"
#include "potrf2_args.h"

namespace dsm {

struct TrtriGen {
  const double* F; const double* W; const double* WT;   // W/WT: the leaf's tiled diagonal-block inverses
  int np, nb, nkc, J;
  int I, c, stored;
  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (I >= nb) return false;
    const int wi = blk_width(np, I);
    const int n1 = BLK / KC, n2 = (I - J - 1) * (BLK / KC), n3 = tri_epilogue_nstages(wi / 32);
    if (c < n1) {                                        // K = J block: X_JJ = W_J
      d.a = WT + (int64_t)J * WBLK_D + c * TILE_D; d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, J * n1 + c, nkc); d.bbytes = TILE_BYTES;
      d.flag0 = nullptr; d.flag1 = nullptr;
    } else if (c < n1 + n2) {
      const int kc = (J + 1) * n1 + (c - n1);
      if (kc / n1 > stored) return false;               // X block of that k-range not stored yet (same CTA)
      d.a = F + tile_off(J, kc, nkc); d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, kc, nkc); d.bbytes = TILE_BYTES;
      d.flag0 = nullptr; d.flag1 = nullptr;
    } else {
      d = tri_epilogue_chunk(W + (int64_t)I * WBLK_D, c - n1 - n2, n3, nullptr);
    }
    if (++c == n1 + n2 + n3) { c = 0; I++; }
    return true;
  }
};

__global__ void __launch_bounds__(NTHREADS, 1) trtri2_kernel(Trtri2Args a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task;
  __shared__ double s_red[16];
  __shared__ double s_al[BLK];
  __shared__ __align__(32) double s_z[NTHREADS / 32][BLK];     // per-warp copy of z_I (prefetched at the start of an iteration)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = 16 * warp;
  Pipe p;
  p.init(smem, a.gerr);
  for (;;) {
    if (tid == 0) s_task = atomicAdd(a.counter, 1);
    __syncthreads();
    const int ti = s_task;
    __syncthreads();
    if (ti >= a.ntasks) return;
    const int2 tk = a.tasks[ti];
    const LeafMeta m = a.meta[tk.x];
    const int J = tk.y, j0 = J * BLK;
    const int wj = blk_width(m.np, J);
    double* F = a.F + m.foff;
    const double* z = a.z + m.voff;
    TrtriGen gen;
    gen.F = F; gen.W = a.W + m.woff; gen.WT = a.WT + m.woff; gen.np = m.np; gen.nb = m.nb; gen.nkc = m.nkc;
    gen.J = J; gen.I = J + 1; gen.c = 0; gen.stored = J;
    double tr = 0.0;
    double al0 = 0.0, al1 = 0.0;       // alpha partial of rows acc_row(0) / acc_row(1) (valid in lanes with t == 0)
    for (int I = J + 1; I < m.nb; I++) {
      const int wi = blk_width(m.np, I), i0 = I * BLK;
      const int nmain = (i0 - j0) / KC;
      Acc2 acc;
      acc2_zero(acc);
      {   // z_I for the fused alpha reduction: latency hidden behind the contraction
        const double4 zv = (4 * lane < wi) ? *reinterpret_cast<const double4*>(z + i0 + 4 * lane) : make_double4(0.0, 0.0, 0.0, 0.0);
        __syncwarp();
        *reinterpret_cast<double4*>(&s_z[warp][4 * lane]) = zv;
        __syncwarp();
      }
      for (int c = 0; c < nmain; c++) {
        if (warp == 0) topup(p, gen, p.q_cons);
        const int st = p.wait();
        if (wi == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0);
        p.release();
      }
      tri_epilogue(p, [&](uint32_t need) { topup(p, gen, need); }, acc, wi / 32, true, -1.0);      // OUT = X_IJ^T  (rows c of block J, cols r of block I)
      acc2_store(acc, F, m.nkc, j0, i0, BLK, wi);
      // fused reductions: ||X_IJ||_F^2 over real rows/cols, alpha_J += X_IJ^T z_I
      double p0 = 0.0, p1 = 0.0;
      const bool row0 = (j0 + acc_row(0)) < m.n, row1 = (j0 + acc_row(1)) < m.n;
#pragma unroll
      for (int n = 0; n < 16; n++) {
        if (8 * n < wi) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int col = i0 + acc_col(n, e);
            if (col < m.n) {
              const double zi = s_z[warp][col - i0];
              const double v0 = acc[0][n][e], v1 = acc[1][n][e];
              if (row0) { tr = fma(v0, v0, tr); p0 = fma(v0, zi, p0); }
              if (row1) { tr = fma(v1, v1, tr); p1 = fma(v1, zi, p1); }
            }
          }
        }
      }
      p0 += __shfl_xor_sync(0xffffffffu, p0, 1); p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
      p1 += __shfl_xor_sync(0xffffffffu, p1, 1); p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
      al0 += p0; al1 += p1;
      __syncthreads();                 // X_IJ^T is in global memory for every warp's rows
      if (warp == 0) { fence_proxy_async(); gen.stored = I; }
    }
    // alpha_J = W_J^T z_J + sum_I X_IJ^T z_I
    if ((lane & 3) == 0) { s_al[acc_row(0)] = al0; s_al[acc_row(1)] = al1; }
    __syncthreads();
    if (tid < wj) {
      const double* WTj = a.WT + m.woff + (int64_t)J * WBLK_D;
      double s = 0.0;
      for (int k = tid; k < wj; k++) s = fma(WTj[widx(tid, k)], z[j0 + k], s);
      a.alpha[m.voff + j0 + tid] = (j0 + tid < m.n) ? s + s_al[tid] : 0.0;
    }
    tr = block_sum(tr, s_red);
    if (tid == 0) a.trpart[a.trpart_off[tk.x] + m.nb + J] = tr;
  }
}

}  // namespace dsm
