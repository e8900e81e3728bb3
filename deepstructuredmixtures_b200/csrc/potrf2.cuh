// Persistent, dependency-driven batched Cholesky (engine v2).
//
// Replaces LAPACK.potrf!('L', F) (gaussianprocess.jl:101) + the forward solve of gaussianprocess.jl:105 for ALL
// experts in ONE launch.  Left-looking by 128-wide block columns; a task is one 128x128 tile:
//   diag (J,J):   C = F_JJ - sum_{K<J} L_JK L_JK^T ; L_JJ = chol(C) ; W_J = L_JJ^-1 ; z_J = W_J (y_J - sum_K L_JK z_K)
//   panel (I,J):  L_IJ = (F_IJ - sum_{K<J} L_IK L_JK^T) W_J^T
// CTAs (one per SM) claim tasks IN ORDER from a global counter.  The host writes the task list in a topological
// order with look-ahead (the diag tile of column J+1 is scheduled right behind the first panel tile of column J)
// and with every expert's columns shifted so that all experts finish together.  A task only ever waits on tasks
// that precede it in the list, and those have been claimed by resident CTAs, so the spin waits cannot deadlock.
// Cross-task ordering is by per-tile flags (release after the tile is in global memory, acquire before the
// producer warp issues the bulk copies of the k-block that reads it).
#pragma once
#include "engine2.cuh"
#include "diag_block.cuh"
#include "args.h"
#include "potrf2_args.h"

namespace dsm {

__device__ __forceinline__ int tile_flag_index(int I, int J) { return I * (I + 1) / 2 + J; }

struct PotrfGen {
  const double* F; const double* z; const double* Wj; const int* flags;
  int nkc, I, J, wj;
  int nc, nmain, nepi, c;                          // stages: F_IJ itself, contraction chunks, W_J (panel tiles only)
  bool diag;
  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (c >= nc + nmain + nepi) return false;
    if (c < nc) {                                  // the tile being updated: 2 column tiles per stage, no dependency
      d.a = F + tile_off(I, J * 8 + 2 * c, nkc); d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, J * 8 + 2 * c + 1, nkc); d.bbytes = TILE_BYTES;
      d.flag0 = nullptr; d.flag1 = nullptr;
    } else if (c < nc + nmain) {
      const int cc = c - nc;
      const int Kb = cc >> 3;                      // 8 chunks per 128-wide k-block
      const bool first = (cc & 7) == 0;
      d.a = F + tile_off(I, cc, nkc); d.abytes = TILE_BYTES;
      if (diag) {             // B operand == A operand; the B part of the stage carries z[16cc .. 16cc+16)
        d.b = z + KC * cc; d.bbytes = KC * 8;
        d.flag0 = first ? flags + tile_flag_index(J, Kb) : nullptr; d.flag1 = nullptr;
      } else {
        d.b = F + tile_off(J, cc, nkc); d.bbytes = TILE_BYTES;
        d.flag0 = first ? flags + tile_flag_index(I, Kb) : nullptr;
        d.flag1 = first ? flags + tile_flag_index(J, Kb) : nullptr;
      }
    } else {
      d = tri_epilogue_chunk(Wj, c - nc - nmain, flags + tile_flag_index(J, J));
    }
    c++;
    return true;
  }
};

__global__ void __launch_bounds__(NTHREADS, 1) potrf2_kernel(Potrf2Args a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_task;
  __shared__ double s_red[16];
  __shared__ double s_v[BLK];
  __shared__ int s_info;
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
    const int4 tk = a.tasks[ti];
    const LeafMeta m = a.meta[tk.x];
    const int I = tk.y, J = tk.z;
    const int i0 = I * BLK, j0 = J * BLK;
    const int wi = blk_width(m.np, I), wj = blk_width(m.np, J);
    double* F = a.F + m.foff;
    int* flags = a.flags + a.flag_off[tk.x];
    double* Wj = a.W + m.woff + (int64_t)J * WBLK_D;
    const int nkc = m.nkc;
    const bool diag = (I == J);
    const bool active = r0 < wi;
    const bool prefactored = (J < a.jstart);     // chol_continue: column already final, diag only rebuilds W
    if (!diag && I < a.jstart) {                  // tile already final
      if (tid == 0) st_release(flags + tile_flag_index(I, J), 1);
      continue;
    }
    PotrfGen gen;
    gen.F = F; gen.z = a.z + m.voff; gen.Wj = Wj; gen.flags = flags; gen.nkc = nkc;
    gen.wj = wj; gen.I = I; gen.J = J; gen.diag = diag; gen.c = 0;
    gen.nmain = (diag && prefactored) ? 0 : j0 / KC;
    gen.nepi = diag ? 0 : tri_epilogue_nstages(wj / 32);
    gen.nc = wj / 32;

    // acc = -F_IJ: the tile arrives through the ring as the first wj/32 stages (two 16-column tiles per stage)
    Acc2 acc;
    acc2_zero(acc);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      if (e < gen.nc) {
        if (warp == 0) topup(p, gen);
        const int st = p.wait();
        if (active) { acc2_sub_tile(acc, p.A(st), r0, 2 * e); acc2_sub_tile(acc, p.B(st), r0, 2 * e + 1); }
        p.release();
      }
    }

    if (!diag) {
      // ---------------- panel tile ----------------
      const int nmain = gen.nmain;
      for (int c = 0; c < nmain; c++) {
        if (warp == 0) topup(p, gen);
        const int st = p.wait();
        if (active) { if (wj == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0); }
        p.release();
      }
      // X = C W_J^T = (-acc) W_J^T, in registers (M = W_J streamed through the ring)
      tri_epilogue(p, gen, acc, wj / 32, active, -1.0);
      acc2_store(acc, F, nkc, i0, j0, wi, wj);
      __syncthreads();
      if (tid == 0) { __threadfence(); st_release(flags + tile_flag_index(I, J), 1); }
      continue;
    }

    // ---------------- diagonal tile ----------------
    double gemv = 0.0;                                  // row r0 + (lane & 15), k-half lane >> 4
    {
      const int nmain = gen.nmain;
      const int ng = min(wj / 32, warp / 2 + 1);        // lower triangle only: columns <= 16*warp + 15
      for (int c = 0; c < nmain; c++) {
        if (warp == 0) topup(p, gen);
        const int st = p.wait();
        if (active) {
          const double* sA = p.A(st);
          switch (ng) {
            case 1: mma_chunk<1>(acc, sA, sA, r0); break;
            case 2: mma_chunk<2>(acc, sA, sA, r0); break;
            case 3: mma_chunk<3>(acc, sA, sA, r0); break;
            default: mma_chunk<4>(acc, sA, sA, r0); break;
          }
          const double* zs = p.B(st);
          const double* ar = sA + r0 + (lane & 15) + (lane >> 4) * 8 * LDS;
#pragma unroll
          for (int k = 0; k < 8; k++) gemv = fma(ar[k * LDS], zs[(lane >> 4) * 8 + k], gemv);
        }
        p.release();
      }
    }
    gemv += __shfl_xor_sync(0xffffffffu, gemv, 16);
    __syncthreads();                                     // every warp is done with the ring: stages become scratch
    double* S = smem;                                    // resident tile [c][LDS] (stages 0..3)
    double* aux = smem + 4 * STAGE_DOUBLES;              // stage 4: scratch
    if (active) {
#pragma unroll
      for (int n = 0; n < 16; n++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = acc_col(n, e);
          if (c < wj) *reinterpret_cast<double2*>(S + c * LDS + acc_row(0)) = make_double2(-acc[0][n][e], -acc[1][n][e]);
        }
      if (lane < 16) s_v[r0 + lane] = gemv;
    }
    __syncthreads();
    {
      const int info = diag_factor_invert(S, wj, aux, !prefactored, &s_info);
      if (tid == 0 && info != 0) atomicCAS(&a.scal[tk.x].info, 0, j0 + info);
    }
    const double* DI = aux;
    if (!prefactored) {
      for (int c = warp; c < wj; c += NTHREADS / 32) {
        double* dst = F + tidx(j0, j0 + c, nkc);
        for (int r = lane; r < wj; r += 32)
          if (r >= c) dst[r] = S[c * LDS + r];
      }
    }
    double ld = 0.0;
    for (int r = tid; r < wj; r += NTHREADS)
      if (j0 + r < m.n) ld += log(S[r * LDS + r]);
    ld = block_sum(ld, s_red);
    // W_J, W_J^T (zero filled, tiled) and the diagonal-block partial of tr(F^-1)
    double* WTj = a.WT + m.woff + (int64_t)J * WBLK_D;
    double tr = 0.0;
    for (int c = warp; c < BLK; c += NTHREADS / 32)
      for (int r = lane; r < BLK; r += 32) {
        double v = 0.0;
        if (r < wj && c < wj && r >= c) v = diag_W(S, DI, r, c);
        Wj[widx(r, c)] = v;
        if (j0 + r < m.n && j0 + c < m.n) tr += v * v;
      }
    for (int r = warp; r < BLK; r += NTHREADS / 32)
      for (int c = lane; c < BLK; c += 32) {
        double v = 0.0;
        if (r < wj && c < wj && r >= c) v = diag_W(S, DI, r, c);
        WTj[widx(c, r)] = v;
      }
    tr = block_sum(tr, s_red);
    // forward solve block: z_J = W_J (y_J - sum_K L_JK z_K)
    if (tid < BLK) s_v[tid] = (tid < wj) ? a.y[m.voff + j0 + tid] - s_v[tid] : 0.0;
    __syncthreads();
    double zz = 0.0;
    {
      const int r = tid >> 1, h = tid & 1;              // 2 threads per row
      double s = 0.0;
      if (r < wj)
        for (int k = h; k <= r; k += 2) s = fma(diag_W(S, DI, r, k), s_v[k], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (r < wj && h == 0) {
        a.z[m.voff + j0 + r] = s;
        if (j0 + r < m.n) zz = s * s;
      }
    }
    zz = block_sum(zz, s_red);
    if (tid == 0) {
      const int64_t po = a.trpart_off[tk.x];
      a.trpart[po + J] = tr;
      a.ldpart[po / 2 + J] = 2.0 * ld;
      a.zzpart[po / 2 + J] = zz;
    }
    fence_proxy_async();                                 // generic writes to the stages precede the next bulk copies
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release(flags + tile_flag_index(J, J), 1); }
  }
}

}  // namespace dsm
