// Block back-substitution  alpha = L^{-T} z  for the fit-only path and for consumers of alpha itself after a gradient
// evaluation (gaussianprocess.jl:105); the forward solve z = L^{-1} y is fused into the diagonal tiles of potrf2.cuh and,
// on the gradient path, alpha = X^T z comes out of trtri3.cuh.
//
//   alpha_J = W_J^T ( z_J - sum_{I > J} L_IJ^T alpha_I ),      J = nb-1 ... 0          (W_J = L_JJ^-1 from the factorisation)
//
// Memory bound (every tile of L is read once).  One task per (expert, block column J); persistent CTAs claim tasks in
// list order (descending J inside an expert), a task streams its column block from the bottom up and acquires the flag of
// alpha_I just before it needs it, so the column blocks of one expert are in flight on many SMs at once and only the last
// block of a task waits for the task right before it.  (The first version walked a whole expert with ONE CTA: 4.7 ms
// on cfg3, bound by single-SM bandwidth.)
#pragma once
#include "pipe.cuh"
#include "args.h"

namespace dsm {

__global__ void __launch_bounds__(NTHREADS) solve3_kernel(SolveArgs a) {
  __shared__ double s_part[NTHREADS / 32][KC];   // per-warp partial of its 16 columns
  __shared__ double s_v[BLK];                    // z_J - s
  __shared__ int s_task;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) s_task = atomicAdd(a.counter, 1);
    __syncthreads();
    const int ti = s_task;
    __syncthreads();
    if (ti >= a.ntasks) return;
    const int2 tk = a.tasks[ti];
    if (a.share != nullptr && a.share[tk.x].x == SHARE_ALIAS) continue;     // block-uniform: alpha of the source expert is used
    const LeafMeta m = a.meta[tk.x];
    const int J = tk.y, j0 = J * BLK, wj = blk_width(m.np, J);
    const double* F = a.F + m.foff;
    const double* al = a.alpha + m.voff;
    int* flags = a.flags + a.flag_off[tk.x];
    // warp w owns the 16 columns [16 w, 16 w + 16) of the block = tile (I, 8 J + w) of every row block I below
    double acc[KC];
#pragma unroll
    for (int k = 0; k < KC; k++) acc[k] = 0.0;
    const bool wact = 16 * warp < wj;
    for (int I = m.nb - 1; I > J; I--) {
      if (lane == 0) {                             // alpha_I is produced by another CTA (bounded spin, like Pipe::wait_flag)
        if (ld_acquire(flags + I) == 0) {
          const long long t0 = clock64();
          while (ld_acquire(flags + I) == 0) {
            if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) { atomicCAS(a.gerr, 0, 7); break; }
            __nanosleep(32);
          }
        }
      }
      __syncwarp();
      if (!wact) continue;
      const int wi = blk_width(m.np, I);
      const double* tile = F + tile_off(I, J * (BLK / KC) + warp, m.nkc);
      // lane holds rows 4 lane .. 4 lane + 3 of the row block
      double4 av = make_double4(0.0, 0.0, 0.0, 0.0);
      if (4 * lane < wi) av = *reinterpret_cast<const double4*>(al + I * BLK + 4 * lane);
      if (4 * lane < wi) {
#pragma unroll
        for (int k = 0; k < KC; k++) {
          const double4 lv = *reinterpret_cast<const double4*>(tile + k * LDS + 4 * lane);
          acc[k] = fma(lv.x, av.x, fma(lv.y, av.y, fma(lv.z, av.z, fma(lv.w, av.w, acc[k]))));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const double s = warp_sum(acc[k]);
      if (lane == 0) s_part[warp][k] = s;
    }
    __syncthreads();
    if (tid < BLK) s_v[tid] = (tid < wj) ? a.z[m.voff + j0 + tid] - s_part[tid >> 4][tid & 15] : 0.0;
    __syncthreads();
    // alpha_J = W_J^T s_v :  alpha[c] = sum_{k >= c} W[k][c] s_v[k];  WT tile layout: WT[c][k] at widx(c, k), c contiguous
    const double* WTd = a.WT + m.woff + (int64_t)J * WBLK_D;
    {
      const int c = tid & (BLK - 1), half = tid >> 7;
      double t = 0.0;
      if (c < wj) {
        const int mid = (c + wj) / 2;
        const int kb = half ? mid : c, ke = half ? wj : mid;
        for (int k = kb; k < ke; k++) t = fma(WTd[widx(c, k)], s_v[k], t);
      }
      __syncthreads();
      if (half) s_v[c] = t;                       // s_v is dead as an input now: reuse it for the upper-half partial
      __syncthreads();
      if (!half && c < wj) a.alpha[m.voff + j0 + c] = (j0 + c < m.n) ? t + s_v[c] : 0.0;
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); st_release(flags + J, 1); }
  }
}

}  // namespace dsm
