// Batched predictive mean / variance of the leaf experts on engine v2 (producer warp + bulk-copy ring).
//
// Replaces prediction(gp, xtest) gaussianprocess.jl:110-137:  mu = m + Knt' alpha ;  V = L \ Knt ;
// Sigma = Ktt - V'V ; diag += exp(2 logNoise).  Only diag(Sigma) is consumed (common.jl:136,147), so the
// T x T matrices Ktt and V'V are never formed:  var_t = k(x_t,x_t) - sum_k V_kt^2 + eta.
//
// Task = (leaf, block Q of 128 routed test points) -- or, in WAVE mode (few test points, so few such tasks), one task
// per (leaf, Q, row block I) with per-block flags: the row blocks of one forward substitution then stream L from HBM on
// many SMs at once (a wavefront: only the last k-block of a task waits for the block above it) instead of one CTA
// walking the whole factor.  Block forward substitution, TRANSPOSED so that a warp owns
// complete rows (= test points) and the multiplication by W_I = L_II^-1 runs in registers:
//   for I = 0..nb-1:   OUT[c][r] = sum_{K<I} V_KQ^T[c][.] L_IK^T[.][r]          (A = V^T tiles, B = L tiles)
//                      V_IQ^T    = (Knt_IQ^T - OUT) W_I^T                        (Knt recomputed from the point tiles)
// V^T of the task lives in a scratch row block with the factor's tile layout, so the next iterations read it back
// with one bulk copy per k-chunk; the consumers signal every stored block to the producer through aux[1].
// The mean and the column sums of squares are accumulated in registers (quad reductions), no shared round trips.
#pragma once
#include "engine2.cuh"
#include "fastexp.cuh"
#include "args.h"

namespace dsm {

constexpr int PRED_DSTAGE = 8;      // dimensions of the point tiles held in static shared memory (else: ring borrowed)

struct Pred3Gen {
  const double* F; const double* W; const double* VT;
  int np, nb, nkc, I, c;
  __device__ __forceinline__ int kneed() const {       // k-block read by the next chunk from V^T, or -1
    if (I >= nb) return -1;
    const int nmain = I * (BLK / KC);
    return c < nmain ? c / (BLK / KC) : -1;
  }
  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (I >= nb) return false;
    const int wi = blk_width(np, I);
    const int nmain = I * (BLK / KC), nepi = tri_epilogue_nstages(wi / 32);
    d.flag0 = nullptr; d.flag1 = nullptr;
    if (c < nmain) {
      d.a = VT + (int64_t)c * TILE_D; d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, c, nkc); d.bbytes = TILE_BYTES;
    } else {
      d = tri_epilogue_chunk(W + (int64_t)I * WBLK_D, c - nmain, nepi, nullptr);
    }
    if (++c == nmain + nepi) { c = 0; I++; }
    return true;
  }
};

__device__ __forceinline__ void predict3_producer(Pipe& p, const PredArgs& a) {
  uint32_t stored_phase = 0, scratch_phase = 0;
  const bool borrow = a.D > PRED_DSTAGE;
  for (;;) {
    int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(a.counter, 1);
    const int ti = __shfl_sync(0xffffffffu, t, 0);
    if (ti >= a.ntasks) break;
    int4 wt = make_int4(0, 0, 0, 0);
    if (a.wave) wt = a.wtasks[ti];
    const int2 tk = a.wave ? make_int2(wt.x, wt.y) : a.tasks[ti];
    const PredLeaf pl = a.pl[tk.x];
    const LeafMeta m = a.meta[pl.slot];
    Pred3Gen gen;
    gen.F = a.F + m.foff; gen.W = a.W + m.woff; gen.np = m.np; gen.nb = a.wave ? wt.z + 1 : m.nb; gen.nkc = m.nkc;
    gen.VT = a.vt_per_cta ? a.VT + (int64_t)blockIdx.x * a.vt_stride : a.VT + pl.vtoff + (int64_t)tk.y * m.nkc * TILE_D;
    gen.I = a.wave ? wt.z : 0; gen.c = 0;
    TaskHdr h; h.kind = 0; h.ti = ti; h.slot = tk.x; h.I = gen.I; h.J = tk.y; h.wi = 0; h.wj = 0; h.n_c = 0; h.n_main = 0; h.pad0 = wt.w;
    int stored = a.wave ? gen.I : 0;             // blocks of V^T in memory (wave mode: ordered by the flags below instead)
    int nsig = a.wave ? 0 : m.nb;                // signals the consumers send for this task (one per iteration)
    bool first = true, allready = false;
    ChunkDesc d;
    if (borrow) {                                // header-only chunk: the ring must be empty whenever the consumers borrow it
      d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
      p.issue(d, &h); first = false;
    }
    for (;;) {
      const int kn = gen.kneed();
      while (kn >= stored) { p.wait_bar(&p.aux[1], stored_phase & 1, 6); stored_phase++; stored++; nsig--; fence_proxy_async(); if (*p.abort) break; }
      if (borrow && gen.c == gen.I * (BLK / KC) && gen.I < gen.nb) {
        // the consumers borrow the ring for the point tiles between the contraction and the epilogue of an iteration
        p.wait_bar(&p.aux[0], scratch_phase & 1, 5); scratch_phase++;
      }
      const int cnext = gen.c, Inext = gen.I;
      if (!gen.next(d)) break;
      if (a.wave && cnext < Inext * (BLK / KC) && (cnext & (BLK / KC - 1)) == 0 && !allready) {
        // block K = cnext / 8 of V^T is written by another CTA: acquire its flag (the last block's flag implies all)
        const int* fl = a.flags + wt.w;
        if (cnext == 0 && Inext > 1) {
          int v = 1;
          if ((threadIdx.x & 31) == 0) v = ld_acquire(fl + Inext - 1);
          allready = __shfl_sync(0xffffffffu, v, 0) != 0;
          if (allready) fence_proxy_async();
        }
        if (!allready) d.flag0 = fl + cnext / (BLK / KC);
      }
      p.issue(d, first ? &h : nullptr); first = false;
      if (*p.abort) break;
    }
    while (nsig > 0 && !*p.abort) { p.wait_bar(&p.aux[1], stored_phase & 1, 6); stored_phase++; nsig--; }
    if (*p.abort) break;
  }
  TaskHdr h; h.kind = -1;
  ChunkDesc d; d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
  p.issue(d, &h);
}

// Kernel values of NG groups of 2 test points (rows rb, rb+1 of the test tile) x 4 consecutive training rows (cb + 16 g .. + 3).
// The exp chains of the SE kernels are long (10 dependent FP64 operations + a table load) and a CTA has only two MMA warps per
// scheduler, so the groups are evaluated NG = 2 at a time: 16 independent chains per thread keep the FP64 pipe busy (measured with
// the per-task trace: 70.7 -> see profiles/ us per 128 x 128 ArdSE tile).  One out-of-line copy per kernel type instead of eight inlined
// ones keeps the kernel inside the instruction cache (383 KB -> 172 KB of SASS, +4.5 % / +10 % on cfg3 / cfg4).
#ifndef DSM_KG_NG
#define DSM_KG_NG 2
#endif
template <int KT, int NG>
__device__ __noinline__ void kernel_groups(int D, const double* sxq, const double* sxi, const double* scf,
                                           const double* sT, double v, int rb, int cb, double (&kv)[NG][2][4]) {
#pragma unroll
  for (int g = 0; g < NG; g++)
#pragma unroll
    for (int mm = 0; mm < 2; mm++)
#pragma unroll
      for (int k = 0; k < 4; k++) kv[g][mm][k] = 0.0;
#pragma unroll 2
  for (int d = 0; d < D; d++) {
    const double2 xq = *reinterpret_cast<const double2*>(sxq + d * BLK + rb);
    const double xqv[2] = {xq.x, xq.y};
    const double cf = scf[d];
#pragma unroll
    for (int g = 0; g < NG; g++) {
      const double2 xi0 = *reinterpret_cast<const double2*>(sxi + d * BLK + cb + 16 * g), xi1 = *reinterpret_cast<const double2*>(sxi + d * BLK + cb + 16 * g + 2);
      const double xiv[4] = {xi0.x, xi0.y, xi1.x, xi1.y};
#pragma unroll
      for (int mm = 0; mm < 2; mm++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (KT == ISO_SE) { const double t = xqv[mm] - xiv[k]; kv[g][mm][k] = fma(t, t, kv[g][mm][k]); }
          else if (KT == ARD_SE) { const double t = xqv[mm] - xiv[k]; kv[g][mm][k] += exp_neg(cf * (t * t), sT); }
          else if (KT == ISO_LINEAR) kv[g][mm][k] = fma(xqv[mm], xiv[k], kv[g][mm][k]);
          else kv[g][mm][k] = fma(cf * xqv[mm], xiv[k], kv[g][mm][k]);
        }
    }
  }
  const double cf0 = scf[0];
#pragma unroll
  for (int g = 0; g < NG; g++)
#pragma unroll
    for (int mm = 0; mm < 2; mm++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (KT == ISO_SE) kv[g][mm][k] = v * exp_neg(cf0 * kv[g][mm][k], sT);
        else if (KT == ARD_SE) kv[g][mm][k] = v * kv[g][mm][k];
        else if (KT == ISO_LINEAR) kv[g][mm][k] = cf0 * kv[g][mm][k];
      }
}

template <int NG>
__device__ __forceinline__ void kernel_groups_any(int ktype, int D, const double* sxq, const double* sxi, const double* scf,
                                                  const double* sT, double v, int rb, int cb, double (&kv)[NG][2][4]) {
  if (ktype == ISO_SE) kernel_groups<ISO_SE, NG>(D, sxq, sxi, scf, sT, v, rb, cb, kv);
  else if (ktype == ARD_SE) kernel_groups<ARD_SE, NG>(D, sxq, sxi, scf, sT, v, rb, cb, kv);
  else if (ktype == ISO_LINEAR) kernel_groups<ISO_LINEAR, NG>(D, sxq, sxi, scf, sT, v, rb, cb, kv);
  else kernel_groups<ARD_LINEAR, NG>(D, sxq, sxi, scf, sT, v, rb, cb, kv);
}

// STRIP mode (at most 16 real test points in the block, i.e. only slab 0 is live): the 128 columns of the 16-row strip are
// split over the eight warps -- warp w contracts columns [16w, 16w+16) -- so that a task streams L at the bulk-copy
// rate instead of being bound by ONE warp issuing 128 DMMAs per chunk.  sacc[mb][n][e]: row 2g+mb, col 16w + 2(2t+e) + n.
__device__ __forceinline__ void strip_chunk(double (&sacc)[2][2][2], const double* __restrict__ sA, const double* __restrict__ sB, int col0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double* pa = sA + t * LDS + 2 * g;
  const double* pb = sB + t * LDS + 2 * g + col0;
#pragma unroll
  for (int ks = 0; ks < KC / 4; ks++) {
    const double2 av = *reinterpret_cast<const double2*>(pa + ks * 4 * LDS);
    const double2 bv = *reinterpret_cast<const double2*>(pb + ks * 4 * LDS);
    dmma884(sacc[0][0][0], sacc[0][0][1], av.x, bv.x);
    dmma884(sacc[1][0][0], sacc[1][0][1], av.y, bv.x);
    dmma884(sacc[0][1][0], sacc[0][1][1], av.x, bv.y);
    dmma884(sacc[1][1][0], sacc[1][1][1], av.y, bv.y);
  }
}

__global__ void __launch_bounds__(NTHREADS_PW, 1) predict3_kernel(PredArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(16) double s_stage[2 * PRED_DSTAGE * BLK + BLK + PRED_DSTAGE];
  __shared__ double sT[EXPTAB_N];
  __shared__ double s_xch[4][8][32];               // strip mode: exchange of the column-split strip (two rounds of 4 warps)
  __shared__ double s_musum[NCONS / 32][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp_slab(), r0 = 16 * slab;
  exptab_load(sT);
  Pipe p;
  p.init(smem, a.gerr);
  if (warp >= NCONS / 32) {
    setmaxnreg_dec<REGS_PRODUCER>();
    if (warp == NCONS / 32) predict3_producer(p, a);
    return;
  }
  setmaxnreg_inc<REGS_CONSUMER>();
  const int D = a.D;
  const bool borrow = D > PRED_DSTAGE;
  // D <= PRED_DSTAGE: both point tiles in static shared memory.  Larger D: x_t stays static only if it fits, so both
  // tiles move into the (idle) ring between the contraction and the epilogue of every iteration.
  double* stage = borrow ? smem : s_stage;
  double* sxq = stage;                   // [D][BLK] test points of block Q
  double* sxi = stage + D * BLK;         // [D][BLK] rows of block I
  double* sal = stage + 2 * D * BLK;     // [BLK] alpha of block I
  double* scf = sal + BLK;               // [D]
  for (;;) {
    int st = p.wait();
    const TaskHdr hd = p.hdr[st];
    if (hd.kind < 0 || *p.abort) return;
    const PredLeaf pl = a.pl[hd.slot];
    const LeafMeta m = a.meta[pl.slot];
    const int Q = hd.J, q0 = Q * BLK;
    const int64_t lda = m.np, ldv = pl.Tp;
    const double* x = a.xg + m.xoff;
    const double* xt = a.xt + pl.xtoff;
    const double* al = a.alpha + m.voff;
    const double* prm = a.prm + m.poff;
    double* VT = a.vt_per_cta ? a.VT + (int64_t)blockIdx.x * a.vt_stride : a.VT + pl.vtoff + (int64_t)Q * m.nkc * TILE_D;
    const int* pidx = a.pidx != nullptr ? a.pidx + pl.ooff + q0 : nullptr;       // device-routed points: gather from xtest
    const int ktype = m.ktype;
    const double v = prm[PRM_V];
    const bool c0ok = q0 + acc_row(0) < pl.T, c1ok = q0 + acc_row(1) < pl.T;
    // slabs without a real test point skip all arithmetic: with a handful of points per expert the task is then bound by
    // the bulk copies of L (HBM), not by 128-wide DMMA tiles that are mostly padding
    const bool active = q0 + r0 < pl.T;
    const bool strip = pl.T - q0 <= 16;              // only slab 0 (= warp 0) holds real test points
    double mu0 = 0.0, mu1 = 0.0, sq0 = 0.0, sq1 = 0.0;
    // optional per-task cycle counts: [0] start, [1] contraction, [2] staging, [3] kernel values, [4] TRSM epilogue, [5] store, [6] end, [7] I count
    long long* trc = (a.trace != nullptr && tid == 0) ? a.trace + (long long)hd.ti * 8 : nullptr;
    long long tc0 = 0, t_con = 0, t_stage = 0, t_eval = 0, t_epi = 0, t_store = 0;
    if (trc) { trc[0] = clock64(); }
    if (!borrow) {
      csync();                                       // previous task's readers are done with the static tiles
      for (int u = tid; u < D * BLK; u += NCONS) {
        const int d = u / BLK, q = u % BLK;
        if (pidx != nullptr) { const int gp = (q0 + q < pl.Tp) ? pidx[q] : -1; sxq[u] = gp >= 0 ? a.xtest[(int64_t)d * a.T_all + gp] : 0.0; }
        else sxq[u] = (q0 + q < pl.Tp) ? xt[(int64_t)d * ldv + q0 + q] : 0.0;
      }
      if (tid < D) scf[tid] = (m.nl > 1) ? prm[PRM_COEF + tid] : prm[PRM_COEF];
    }
    if (borrow) p.release();                         // header-only chunk
    const int g8 = lane >> 2, t4 = lane & 3;
    const int rb = r0 + 2 * g8;                      // this thread's two test points (rows of OUT)
    const int I_begin = a.wave ? hd.I : 0, I_end = a.wave ? hd.I + 1 : m.nb;
    for (int I = I_begin; I < I_end; I++) {
      const int wi = blk_width(m.np, I), i0 = I * BLK;
      Acc2 acc;
      acc2_zero(acc);
      double sacc[2][2][2];
#pragma unroll
      for (int mm = 0; mm < 2; mm++) { sacc[mm][0][0] = 0.0; sacc[mm][0][1] = 0.0; sacc[mm][1][0] = 0.0; sacc[mm][1][1] = 0.0; }
      const bool wact = strip && 16 * warp < wi;     // strip mode: this warp's 16 columns exist in block I
      if (trc) tc0 = clock64();
      for (int c = 0; c < I * (BLK / KC); c++) {     // (I == 0 has no contraction: the header chunk is its first epilogue stage)
        st = p.wait();                                 // (idempotent for the header chunk, which is still unreleased)
        if (strip) { if (wact) strip_chunk(sacc, p.A(st), p.B(st), 16 * warp); }
        else if (active) { if (wi == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0); }
        p.release();
      }
      // stage the point tiles of block I (and, when borrowing the ring, of the test block too)
      if (trc) { const long long t1 = clock64(); t_con += t1 - tc0; tc0 = t1; }
      csync();
      if (borrow) {
        for (int u = tid; u < D * BLK; u += NCONS) {
          const int d = u / BLK, q = u % BLK;
          if (pidx != nullptr) { const int gp = (q0 + q < pl.Tp) ? pidx[q] : -1; sxq[u] = gp >= 0 ? a.xtest[(int64_t)d * a.T_all + gp] : 0.0; }
          else sxq[u] = (q0 + q < pl.Tp) ? xt[(int64_t)d * ldv + q0 + q] : 0.0;
        }
        if (tid < D) scf[tid] = (m.nl > 1) ? prm[PRM_COEF + tid] : prm[PRM_COEF];
      }
      for (int u = tid; u < D * BLK; u += NCONS) {
        const int d = u / BLK, q = u % BLK;
        sxi[u] = (q < wi) ? x[(int64_t)d * lda + i0 + q] : 0.0;
      }
      if (tid < BLK) sal[tid] = (tid < wi && i0 + tid < m.n) ? al[i0 + tid] : 0.0;
      csync();
      if (trc) { const long long t1 = clock64(); t_stage += t1 - tc0; tc0 = t1; }
      // OUT = Knt_IQ^T - OUT ; mean partial sum_r Knt[r][c] alpha[r].  Groups of 2 test points x 4 consecutive rows r.
      if (strip) {
        if (wact) {
          const int rbs = 2 * g8, cb = 16 * warp + 4 * t4;
          double kv[1][2][4];
          kernel_groups_any<1>(ktype, D, sxq, sxi, scf, sT, v, rbs, cb, kv);
          const double2 al0 = *reinterpret_cast<const double2*>(sal + cb), al1 = *reinterpret_cast<const double2*>(sal + cb + 2);
          const double alv[4] = {al0.x, al0.y, al1.x, al1.y};
#pragma unroll
          for (int mm = 0; mm < 2; mm++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
              double kk = kv[0][mm][k];
              if (!(q0 + rbs + mm < pl.T && i0 + cb + k < m.n)) kk = 0.0;
              if (mm) mu1 = fma(kk, alv[k], mu1); else mu0 = fma(kk, alv[k], mu0);
              sacc[mm][k & 1][k >> 1] = kk - sacc[mm][k & 1][k >> 1];
            }
        }
        // gather the strip in warp 0's accumulator fragments: same lane, tiles 2 cg + n of the interleaved mapping
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          if ((warp >> 2) == hh) {
#pragma unroll
            for (int mm = 0; mm < 2; mm++)
#pragma unroll
              for (int n = 0; n < 2; n++)
#pragma unroll
                for (int e = 0; e < 2; e++) s_xch[warp & 3][mm * 4 + n * 2 + e][lane] = sacc[mm][n][e];
          }
          csync();
          if (warp == 0) {
#pragma unroll
            for (int cg = 0; cg < 4; cg++)
#pragma unroll
              for (int mm = 0; mm < 2; mm++)
#pragma unroll
                for (int n = 0; n < 2; n++)
#pragma unroll
                  for (int e = 0; e < 2; e++) acc[mm][2 * (4 * hh + cg) + n][e] = s_xch[cg][mm * 4 + n * 2 + e][lane];
          }
          csync();
        }
      } else {
#pragma unroll
      for (int nbp2 = 0; nbp2 < 8 / DSM_KG_NG; nbp2++) {
        if (active && 16 * DSM_KG_NG * nbp2 < wi) {     // wi is a multiple of 64: every 16-column group of the set exists
          const int cb0 = 16 * DSM_KG_NG * nbp2 + 4 * t4;
          double kv[DSM_KG_NG][2][4];
          kernel_groups_any<DSM_KG_NG>(ktype, D, sxq, sxi, scf, sT, v, rb, cb0, kv);
#pragma unroll
          for (int g = 0; g < DSM_KG_NG; g++) {
            const int nbp = DSM_KG_NG * nbp2 + g, cb = cb0 + 16 * g;
            const double2 al0 = *reinterpret_cast<const double2*>(sal + cb), al1 = *reinterpret_cast<const double2*>(sal + cb + 2);
            const double alv[4] = {al0.x, al0.y, al1.x, al1.y};
#pragma unroll
            for (int mm = 0; mm < 2; mm++)
#pragma unroll
              for (int k = 0; k < 4; k++) {
                double kk = kv[g][mm][k];
                if (!((mm ? c1ok : c0ok) && i0 + cb + k < m.n)) kk = 0.0;
                if (mm) mu1 = fma(kk, alv[k], mu1); else mu0 = fma(kk, alv[k], mu0);
                acc[mm][2 * nbp + (k & 1)][k >> 1] = kk - acc[mm][2 * nbp + (k & 1)][k >> 1];
              }
          }
        }
      }
      }
      if (borrow) {                                    // hand the ring back before the epilogue stages arrive
        fence_proxy_async();
        csync();
        if (tid == 0) mbar_arrive(&p.aux[0]);
      }
      if (trc) { const long long t1 = clock64(); t_eval += t1 - tc0; tc0 = t1; }
      tri_epilogue(p, acc, wi / 32, active, 1.0);     // V_IQ^T = S^T W_I^T
      if (trc) { const long long t1 = clock64(); t_epi += t1 - tc0; tc0 = t1; }
      if (active) acc2_store(acc, VT, m.nkc, 0, i0, BLK, wi);
      fence_proxy_async();
      csync();
      if (tid == 0) {
        if (a.wave) { __threadfence(); st_release(a.flags + hd.pad0 + I, 1); }     // other CTAs read this block of V^T
        else mbar_arrive(&p.aux[1]);
      }
      if (trc) { const long long t1 = clock64(); t_store += t1 - tc0; tc0 = t1; }
#pragma unroll
      for (int nb = 0; nb < 16; nb++)
        if (8 * nb < wi) {
#pragma unroll
          for (int e = 0; e < 2; e++) { sq0 = fma(acc[0][nb][e], acc[0][nb][e], sq0); sq1 = fma(acc[1][nb][e], acc[1][nb][e], sq1); }
        }
    }
    if (trc) { trc[1] = t_con; trc[2] = t_stage; trc[3] = t_eval; trc[4] = t_epi; trc[5] = t_store; trc[6] = clock64(); trc[7] = I_end - I_begin; }
    // finish: mu = m + Knt' alpha ; var = k(x_t, x_t) - sum V^2 + eta      (gaussianprocess.jl:117-126)
    mu0 += __shfl_xor_sync(0xffffffffu, mu0, 1); mu0 += __shfl_xor_sync(0xffffffffu, mu0, 2);
    mu1 += __shfl_xor_sync(0xffffffffu, mu1, 1); mu1 += __shfl_xor_sync(0xffffffffu, mu1, 2);
    sq0 += __shfl_xor_sync(0xffffffffu, sq0, 1); sq0 += __shfl_xor_sync(0xffffffffu, sq0, 2);
    sq1 += __shfl_xor_sync(0xffffffffu, sq1, 1); sq1 += __shfl_xor_sync(0xffffffffu, sq1, 2);
    if (strip) {                                     // the mean partials of the strip live in all eight warps
      if ((lane & 3) == 0) { s_musum[warp][2 * g8] = mu0; s_musum[warp][2 * g8 + 1] = mu1; }
      csync();
      if (warp == 0) {
        mu0 = 0.0; mu1 = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < NCONS / 32; w8++) { mu0 += s_musum[w8][2 * g8]; mu1 += s_musum[w8][2 * g8 + 1]; }
      }
      csync();
    }
    if (a.wave) {                                    // per-task partials; predict_reduce_kernel finishes
      if ((lane & 3) == 0 && active) {
        double* pp = a.part + (int64_t)(hd.pad0 + hd.I) * 2 * BLK;
        *reinterpret_cast<double2*>(pp + acc_row(0)) = make_double2(mu0, mu1);
        *reinterpret_cast<double2*>(pp + BLK + acc_row(0)) = make_double2(sq0, sq1);
      }
      continue;
    }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int mb = 0; mb < 2; mb++) {
        const int c = acc_row(mb);
        if (q0 + c < pl.T) {
          double ktt;
          if (ktype == ISO_SE) ktt = v;
          else if (ktype == ARD_SE) ktt = v * (double)D;
          else {
            ktt = 0.0;
            for (int d = 0; d < D; d++) {
              const double xv = pidx != nullptr ? a.xtest[(int64_t)d * a.T_all + pidx[c]] : xt[(int64_t)d * ldv + q0 + c];
              const double cf = (ktype == ISO_LINEAR) ? prm[PRM_COEF] : prm[PRM_COEF + d];
              ktt = fma(cf * xv, xv, ktt);
            }
          }
          a.mu[pl.ooff + q0 + c] = a.leaf_mean[m.leaf] + (mb ? mu1 : mu0);
          a.var[pl.ooff + q0 + c] = ktt - (mb ? sq1 : sq0) + prm[PRM_ETA];
        }
      }
    }
  }
}

// Wave mode: mu = m + sum_I partial ; var = k(x_t, x_t) - sum_I partial + eta, summed in block order (deterministic).
__global__ void __launch_bounds__(BLK) predict_reduce_kernel(PredArgs a) {
  const int4 wc = a.wcols[blockIdx.x];
  const PredLeaf pl = a.pl[wc.x];
  const LeafMeta m = a.meta[pl.slot];
  const int q0 = wc.y * BLK, c = threadIdx.x;
  if (q0 + c >= pl.T) return;
  double mu = 0.0, sq = 0.0;
  for (int I = 0; I < wc.w; I++) {
    const double* pp = a.part + (int64_t)(wc.z + I) * 2 * BLK;
    mu += pp[c]; sq += pp[BLK + c];
  }
  const double* prm = a.prm + m.poff;
  const double* xt = a.xt + pl.xtoff;
  const double v = prm[PRM_V];
  double ktt;
  if (m.ktype == ISO_SE) ktt = v;
  else if (m.ktype == ARD_SE) ktt = v * (double)a.D;
  else {
    ktt = 0.0;
    for (int d = 0; d < a.D; d++) {
      const double xv = xt[(int64_t)d * pl.Tp + q0 + c];
      const double cf = (m.ktype == ISO_LINEAR) ? prm[PRM_COEF] : prm[PRM_COEF + d];
      ktt = fma(cf * xv, xv, ktt);
    }
  }
  a.mu[pl.ooff + q0 + c] = a.leaf_mean[m.leaf] + mu;
  a.var[pl.ooff + q0 + c] = ktt - sq + prm[PRM_ETA];
}

}  // namespace dsm
