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
// k-block written by the previous iteration (signalled through the aux[1] barrier).
#pragma once
#include "engine2.cuh"
#include "args.h"
#include "potrf2_args.h"

namespace dsm {

struct TrtriGen {
  const double* F; const double* W; const double* WT;   // W/WT: the leaf's tiled diagonal-block inverses
  int np, nb, nkc, J;
  int I, c;
  // kneed: the k-block the next chunk reads from the X^T rows (-1: none); the caller makes sure it has been stored
  __device__ __forceinline__ int kneed() const {
    const int n1 = BLK / KC, n2 = (I - J - 1) * n1;
    if (I >= nb || c < n1 || c >= n1 + n2) return -1;
    return ((J + 1) * n1 + (c - n1)) / n1;
  }
  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (I >= nb) return false;
    const int wi = blk_width(np, I);
    const int n1 = BLK / KC, n2 = (I - J - 1) * (BLK / KC), n3 = tri_epilogue_nstages(wi / 32);
    d.flag0 = nullptr; d.flag1 = nullptr;
    if (c < n1) {                                        // K = J block: X_JJ = W_J
      d.a = WT + (int64_t)J * WBLK_D + c * TILE_D; d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, J * n1 + c, nkc); d.bbytes = TILE_BYTES;
    } else if (c < n1 + n2) {
      const int kc = (J + 1) * n1 + (c - n1);
      d.a = F + tile_off(J, kc, nkc); d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, kc, nkc); d.bbytes = TILE_BYTES;
    } else {
      d = tri_epilogue_chunk(W + (int64_t)I * WBLK_D, c - n1 - n2, n3, nullptr);
    }
    if (++c == n1 + n2 + n3) { c = 0; I++; }
    return true;
  }
};

// Producer warp.  X_IJ^T of iteration I is written by this CTA's consumers and read back (as the A operand) from the
// next iteration on: the consumers signal every stored block through aux[1]; the producer waits for block K before it
// issues the first chunk that reads it.  (The consumers can be at most one signal ahead of the producer's count --
// finishing iteration I needs a chunk that the producer only issues after it has seen block I-1 -- so phase parity
// is unambiguous.)
__device__ __forceinline__ void trtri2_producer(Pipe& p, const Trtri2Args& a) {
  uint32_t stored_phase = 0;
  for (;;) {
    int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(a.counter, 1);
    const int ti = __shfl_sync(0xffffffffu, t, 0);
    if (ti >= a.ntasks) break;
    const int2 tk = a.tasks[ti];
    const LeafMeta m = a.meta[tk.x];
    TrtriGen gen;
    gen.F = a.F + m.foff; gen.W = a.W + m.woff; gen.WT = a.WT + m.woff; gen.np = m.np; gen.nb = m.nb; gen.nkc = m.nkc;
    gen.J = tk.y; gen.I = tk.y + 1; gen.c = 0;
    TaskHdr h; h.kind = 0; h.ti = ti; h.slot = tk.x; h.I = 0; h.J = tk.y; h.wi = 0; h.wj = blk_width(m.np, tk.y); h.n_c = 0; h.n_main = 0;
    int stored = tk.y;
    int nsig = m.nb - 1 - tk.y;                            // signals the consumers will send for this task
    bool first = true;
    ChunkDesc d;
    if (gen.I >= gen.nb) {                                 // last block column: no contraction, header only
      d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
      p.issue(d, &h);
      continue;
    }
    for (;;) {
      const int kn = gen.kneed();
      while (kn > stored) { p.wait_bar(&p.aux[1], stored_phase & 1, 6); stored_phase++; stored++; nsig--; fence_proxy_async(); }
      if (!gen.next(d)) break;
      p.issue(d, first ? &h : nullptr); first = false;
      if (*p.abort) break;
    }
    while (nsig > 0) { p.wait_bar(&p.aux[1], stored_phase & 1, 6); stored_phase++; nsig--; }    // drain: keep the phase count exact
    if (*p.abort) break;
  }
  TaskHdr h; h.kind = -1;
  ChunkDesc d; d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
  p.issue(d, &h);
}

__global__ void __launch_bounds__(NTHREADS_PW, 1) trtri2_kernel(Trtri2Args a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_red[16];
  __shared__ double s_al[BLK];
  __shared__ __align__(32) double s_z[NCONS / 32][BLK];     // per-warp copy of z_I (prefetched at the start of an iteration)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = 16 * warp;
  Pipe p;
  p.init(smem, a.gerr);
  if (warp >= NCONS / 32) {                          // producer warpgroup: one working warp, three that only donate registers
    setmaxnreg_dec<REGS_PRODUCER>();
    if (warp == NCONS / 32) trtri2_producer(p, a);
    return;
  }
  setmaxnreg_inc<REGS_CONSUMER>();
  for (;;) {
    // header of the next task = header of its first chunk (peeked; the chunk itself is consumed by the loops below)
    const int st0 = p.wait();
    const TaskHdr hd = p.hdr[st0];
    if (hd.kind < 0 || *p.abort) return;
    const LeafMeta m = a.meta[hd.slot];
    const int J = hd.J, j0 = J * BLK;
    const int wj = hd.wj;
    double* F = a.F + m.foff;
    const double* z = a.z + m.voff;
    if (J + 1 >= m.nb) p.release();    // header-only chunk of the last block column
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
        const int st = p.wait();
        if (wi == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0);
        p.release();
      }
      tri_epilogue(p, acc, wi / 32, true, -1.0);      // OUT = X_IJ^T  (rows c of block J, cols r of block I)
      acc2_store(acc, F, m.nkc, j0, i0, BLK, wi);
      fence_proxy_async();             // this thread's generic stores precede the bulk copies that read them back
      csync();                         // X_IJ^T is in global memory for every warp's rows
      if (tid == 0) mbar_arrive(&p.aux[1]);
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
    }
    // alpha_J = W_J^T z_J + sum_I X_IJ^T z_I
    if ((lane & 3) == 0) { s_al[acc_row(0)] = al0; s_al[acc_row(1)] = al1; }
    csync();
    if (tid < wj) {
      const double* WTj = a.WT + m.woff + (int64_t)J * WBLK_D;
      double s = 0.0;
      for (int k = tid; k < wj; k++) s = fma(WTj[widx(tid, k)], z[j0 + k], s);
      a.alpha[m.voff + j0 + tid] = (j0 + tid < m.n) ? s + s_al[tid] : 0.0;
    }
    tr = block_sum_c(tr, s_red);
    if (tid == 0) a.trpart[a.trpart_off[hd.slot] + m.nb + J] = tr;
  }
}

}  // namespace dsm
