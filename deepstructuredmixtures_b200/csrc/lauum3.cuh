// LAUUM tiles with fused gradient traces on engine v2 (producer warp + bulk-copy ring).
//
// Length-scale terms that need F^-1 element-wise (IsoSE; ArdSE / ArdLinear in mathematical mode; kernels.jl:85-99,
// 146-164, 234-246) use   F^-1_IJ = sum_{K >= I} X_KI^T X_KJ   (X = L^-1, stored transposed in the strict upper block
// triangle of the factor by trtri3; X_II^T = W_I^T) followed by an epilogue that recomputes dK/dlog l_h from the point
// tiles and reduces   sum_ij (a_i a_j - F^-1_ij) dK_ij   per tile -- the n^3 GEMM per hyper-parameter of the reference
// becomes one n^3/3 contraction for all of them.  Task = tile (I, J), J <= I, no cross-task dependency.
// The point tiles are staged in a small static shared buffer when D <= LAUUM_DSTAGE (the ring keeps prefetching the
// next task during the epilogue); for larger D the epilogue borrows the ring as scratch and the producer waits (aux[0]).
#pragma once
#include "engine2.cuh"
#include "fastexp.cuh"
#include "args.h"

namespace dsm {

constexpr int LAUUM_DSTAGE = 8;

// sum_ij M_ij dK_ij / dlog l_h for NG groups of 2 rows x 4 consecutive columns (rows rb, rb+1; columns cb + 16 g .. + 3) and the
// hyper-parameters h0 .. h0+3, added to g[0..3].  M = sym * (alpha_i alpha_j - F^-1_ij) comes from the accumulators.
// Out of line and NG = 2 groups at a time for the same reasons as predict3's kernel_groups: one copy per kernel type (instruction
// cache), 16 independent exp chains per thread (two MMA warps per scheduler cannot hide a 10-operation FP64 chain otherwise).
template <int KT, int NG>
__device__ __noinline__ void lauum_groups(int D, int h0, int nl, const double* sxi, const double* sxj, const double* scf,
                                          const double* sT, double v, int rb, int cb, const double (&mij)[NG][2][4], double (&g)[4]) {
  if (KT == ISO_SE || KT == ISO_LINEAR) {
    double r2[NG][2][4];
#pragma unroll
    for (int q = 0; q < NG; q++)
#pragma unroll
      for (int mm = 0; mm < 2; mm++)
#pragma unroll
        for (int k = 0; k < 4; k++) r2[q][mm][k] = 0.0;
#pragma unroll 2
    for (int d = 0; d < D; d++) {
      const double2 xi = *reinterpret_cast<const double2*>(sxi + d * BLK + rb);
      const double xiv[2] = {xi.x, xi.y};
#pragma unroll
      for (int q = 0; q < NG; q++) {
        const double2 xj0 = *reinterpret_cast<const double2*>(sxj + d * BLK + cb + 16 * q), xj1 = *reinterpret_cast<const double2*>(sxj + d * BLK + cb + 16 * q + 2);
        const double xjv[4] = {xj0.x, xj0.y, xj1.x, xj1.y};
#pragma unroll
        for (int mm = 0; mm < 2; mm++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (KT == ISO_SE) { const double tt = xiv[mm] - xjv[k]; r2[q][mm][k] = fma(tt, tt, r2[q][mm][k]); }
            else r2[q][mm][k] = fma(xiv[mm], xjv[k], r2[q][mm][k]);
          }
      }
    }
    double gs[NG] = {};
#pragma unroll
    for (int q = 0; q < NG; q++)
#pragma unroll
      for (int mm = 0; mm < 2; mm++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (KT == ISO_SE) {
            const double u = scf[0] * r2[q][mm][k];                                  // -0.5 r2 / l^2
            gs[q] = fma(mij[q][mm][k], v * exp_neg(u, sT) * (-2.0 * u), gs[q]);       // K * r2 / l^2
          } else {
            gs[q] = fma(mij[q][mm][k], -2.0 * scf[0] * r2[q][mm][k], gs[q]);
          }
        }
#pragma unroll
    for (int q = 0; q < NG; q++) g[0] += gs[q];
  } else {
    for (int hh = 0; hh < 4 && h0 + hh < nl; hh++) {
      const int d = h0 + hh;
      const double cf = scf[d];
      const double2 xi = *reinterpret_cast<const double2*>(sxi + d * BLK + rb);
      const double xiv[2] = {xi.x, xi.y};
      double gs[NG] = {};
#pragma unroll
      for (int q = 0; q < NG; q++) {
        const double2 xj0 = *reinterpret_cast<const double2*>(sxj + d * BLK + cb + 16 * q), xj1 = *reinterpret_cast<const double2*>(sxj + d * BLK + cb + 16 * q + 2);
        const double xjv[4] = {xj0.x, xj0.y, xj1.x, xj1.y};
#pragma unroll
        for (int mm = 0; mm < 2; mm++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            if (KT == ARD_SE) {
              const double tt = xiv[mm] - xjv[k];
              const double u = cf * (tt * tt);
              gs[q] = fma(mij[q][mm][k], v * exp_neg(u, sT) * (-2.0 * u), gs[q]);
            } else {
              gs[q] = fma(mij[q][mm][k], -2.0 * cf * xiv[mm] * xjv[k], gs[q]);
            }
          }
      }
#pragma unroll
      for (int q = 0; q < NG; q++) g[hh] += gs[q];
    }
  }
}

struct Lauum3Gen {
  const double* A0; const double* B0;   // K = I block operands (WT_I tiles / X_IJ^T tiles)
  const double* A1; const double* B1;   // remaining k range: tiles (I, kc), (J, kc)
  int n0, n1, c;
  TaskHdr h;
  __device__ __forceinline__ void load(const LauumArgs& a, int ti) {
    const int4 tk = a.tasks[ti];
    const LeafMeta m = a.meta[tk.x];
    const int I = tk.y, J = tk.z;
    const double* F = a.F + m.foff;
    const double* WTi = a.WT + m.woff + (int64_t)I * WBLK_D;
    const int wi = blk_width(m.np, I), i0 = I * BLK, k1 = i0 + wi;
    A0 = WTi; B0 = (I == J) ? WTi : F + tile_off(J, i0 / KC, m.nkc);
    A1 = F + tile_off(I, k1 / KC, m.nkc); B1 = F + tile_off(J, k1 / KC, m.nkc);
    n0 = wi / KC; n1 = (m.np - k1) / KC; c = 0;
    h.kind = 0; h.ti = ti; h.slot = tk.x; h.I = I; h.J = J; h.wi = wi; h.wj = blk_width(m.np, J);
    h.n_c = 0; h.n_main = n0 + n1; h.pad0 = tk.w;
    if (a.pre != nullptr && a.pre_base[tk.x] >= 0) {
      // the tile arrives ready-made (-F^-1_IJ): two 16-column tiles per stage, no contraction
      A0 = a.pre + (a.pre_base[tk.x] + tk.w) * (int64_t)WBLK_D;
      n0 = 0; n1 = 0; h.n_main = 0; h.n_c = h.wj / 32;
    }
  }
  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (c >= n0 + n1 + h.n_c) return false;
    d.flag0 = nullptr; d.flag1 = nullptr; d.abytes = TILE_BYTES; d.bbytes = TILE_BYTES;
    if (h.n_c > 0) { d.a = A0 + (int64_t)(2 * c) * TILE_D; d.b = A0 + (int64_t)(2 * c + 1) * TILE_D; }
    else if (c < n0) { d.a = A0 + (int64_t)c * TILE_D; d.b = B0 + (int64_t)c * TILE_D; }
    else { d.a = A1 + (int64_t)(c - n0) * TILE_D; d.b = B1 + (int64_t)(c - n0) * TILE_D; }
    c++;
    return true;
  }
};

__device__ __forceinline__ void lauum3_producer(Pipe& p, const LauumArgs& a) {
  Lauum3Gen gen;
  uint32_t scratch_phase = 0;
  const bool borrow = a.D > LAUUM_DSTAGE;
  for (;;) {
    int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(a.counter, 1);
    const int ti = __shfl_sync(0xffffffffu, t, 0);
    if (ti >= a.ntasks) break;
    if (a.mask != nullptr && a.mask[a.tasks[ti].x] == 0) continue;      // expert without a gradient request
    gen.load(a, ti);
    ChunkDesc d;
    bool first = true;
    while (gen.next(d)) { p.issue(d, first ? &gen.h : nullptr); first = false; }
    if (borrow) { p.wait_bar(&p.aux[0], scratch_phase & 1, 5); scratch_phase++; }
    if (*p.abort) break;
  }
  TaskHdr h; h.kind = -1;
  ChunkDesc d; d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
  p.issue(d, &h);
}

__global__ void __launch_bounds__(NTHREADS_PW, 1) lauum3_kernel(LauumArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_red[16];
  __shared__ __align__(16) double s_stage[2 * LAUUM_DSTAGE * BLK + 2 * BLK + LAUUM_DSTAGE];
  __shared__ double sT[EXPTAB_N];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp_slab(), r0 = 16 * slab;
  exptab_load(sT);                                  // visible after the barrier inside Pipe::init
  Pipe p;
  p.init(smem, a.gerr);
  if (warp >= NCONS / 32) {
    setmaxnreg_dec<REGS_PRODUCER>();
    if (warp == NCONS / 32) lauum3_producer(p, a);
    return;
  }
  setmaxnreg_inc<REGS_CONSUMER>();
  const int D = a.D;
  const bool borrow = D > LAUUM_DSTAGE;
  double* stage = borrow ? smem : s_stage;
  double* sxi = stage;                   // [D][BLK]
  double* sxj = stage + D * BLK;         // [D][BLK]
  double* sai = stage + 2 * D * BLK;     // [BLK]
  double* saj = sai + BLK;
  double* scf = saj + BLK;               // [D]
  for (;;) {
    int st = p.wait();
    const TaskHdr hd = p.hdr[st];
    if (hd.kind < 0 || *p.abort) return;
    const LeafMeta m = a.meta[hd.slot];
    const int I = hd.I, J = hd.J, wi = hd.wi, wj = hd.wj, i0 = I * BLK, j0 = J * BLK;
    const bool active = r0 < wi;
    Acc2 acc;
    acc2_zero(acc);
#pragma unroll
    for (int e = 0; e < 4; e++) {                      // ready-made tile: acc = -(-F^-1_IJ)
      if (e < hd.n_c) {
        if (e > 0) st = p.wait();
        if (active) { acc2_sub_tile(acc, p.A(st), r0, 2 * e); acc2_sub_tile(acc, p.B(st), r0, 2 * e + 1); }
        p.release();
      }
    }
    for (int c = 0; c < hd.n_main; c++) {
      if (c > 0) st = p.wait();
      // K = I block: the A operand is W_I^T (zero for k < row): chunk c is all zero for slabs > c
      if (active && (c >= wi / KC || c >= slab)) { if (wj == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0); }
      p.release();
    }
    // ---- fused epilogue -------------------------------------------------------------------
    const double* prm = a.prm + m.poff;
    const double* x = a.xg + m.xoff;
    const double* al = a.alpha + m.voff;
    const int nl = m.nl, ktype = m.ktype;
    const int64_t lda = m.np;
    csync();                                         // previous task's epilogue reads are over (and, when borrowing, the ring is idle)
    for (int u = tid; u < D * BLK; u += NCONS) {
      const int d = u / BLK, q = u % BLK;
      sxi[u] = (q < wi) ? x[(int64_t)d * lda + i0 + q] : 0.0;
      sxj[u] = (q < wj) ? x[(int64_t)d * lda + j0 + q] : 0.0;
    }
    if (tid < BLK) { sai[tid] = (tid < wi) ? al[i0 + tid] : 0.0; saj[tid] = (tid < wj) ? al[j0 + tid] : 0.0; }
    if (tid < D) scf[tid] = (nl > 1) ? prm[PRM_COEF + tid] : prm[PRM_COEF];
    csync();
    const double v = prm[PRM_V];
    const double sym = (I == J) ? 1.0 : 2.0;
    // The thread's 64 elements are processed in 8 groups of 2 rows x 4 CONSECUTIVE columns (interleaved fragment
    // mapping: tiles 2p / 2p+1 hold the even / odd columns of 16-column group p), so one group needs one double2 of
    // x_i and two double2 of x_j per dimension and gives 8 independent chains to the FP64 pipe.
    const int g8 = lane >> 2, t4 = lane & 3;
    const int rb = r0 + 2 * g8;
    for (int h0 = 0; h0 < nl; h0 += 4) {
      double g[4] = {0.0, 0.0, 0.0, 0.0};
      if (active) {
#pragma unroll
        for (int nbp2 = 0; nbp2 < 4; nbp2++) {
          if (32 * nbp2 < wj) {                         // wj is a multiple of 64: both 16-column groups of the pair exist
            const int cb0 = 32 * nbp2 + 4 * t4;
            double mij[2][2][4];
            const double2 ai = *reinterpret_cast<const double2*>(sai + rb);
            const double aiv[2] = {ai.x, ai.y};
#pragma unroll
            for (int q = 0; q < 2; q++) {
              const int nbp = 2 * nbp2 + q, cb = cb0 + 16 * q;
              const double2 aj0 = *reinterpret_cast<const double2*>(saj + cb), aj1 = *reinterpret_cast<const double2*>(saj + cb + 2);
              const double ajv[4] = {aj0.x, aj0.y, aj1.x, aj1.y};
#pragma unroll
              for (int mm = 0; mm < 2; mm++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const bool ok = (i0 + rb + mm < m.n) && (j0 + cb + k < m.n);
                  mij[q][mm][k] = ok ? sym * (aiv[mm] * ajv[k] - acc[mm][2 * nbp + (k & 1)][k >> 1]) : 0.0;
                }
            }
            if (ktype == ISO_SE) lauum_groups<ISO_SE, 2>(D, h0, nl, sxi, sxj, scf, sT, v, rb, cb0, mij, g);
            else if (ktype == ISO_LINEAR) lauum_groups<ISO_LINEAR, 2>(D, h0, nl, sxi, sxj, scf, sT, v, rb, cb0, mij, g);
            else if (ktype == ARD_SE) lauum_groups<ARD_SE, 2>(D, h0, nl, sxi, sxj, scf, sT, v, rb, cb0, mij, g);
            else lauum_groups<ARD_LINEAR, 2>(D, h0, nl, sxi, sxj, scf, sT, v, rb, cb0, mij, g);
          }
        }
      }
      for (int hh = 0; hh < 4 && h0 + hh < nl; hh++) {
        const double s = block_sum_c(g[hh], s_red);
        if (tid == 0) a.gpart[a.gpart_off[hd.slot] + (int64_t)hd.pad0 * nl + h0 + hh] = s;
      }
    }
    if (borrow) {                                    // hand the ring back to the producer
      fence_proxy_async();
      csync();
      if (tid == 0) mbar_arrive(&p.aux[0]);
    }
  }
}

}  // namespace dsm
