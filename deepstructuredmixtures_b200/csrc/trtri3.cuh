// Batched triangular inverse X = L^-1 as a dependency-driven TILE pipeline (engine v2), with fused partials of
// tr(F^-1) = ||X||_F^2  and  alpha = X^T z.
//
// Replaces the n-RHS ldiv!(cK, -I) of ααinvcK! (gaussianprocess.jl:219-226) and the backward solve of
// gaussianprocess.jl:105.  Task = one 128 x 128 tile (I, J), I > J:
//     S    = sum_{K=J}^{I-1} L_IK X_KJ            (K = J uses X_JJ = W_J from the factorisation)
//     X_IJ = -W_I S
// X_IJ depends on every X_KJ above it in its block column, so a column is a chain -- but only the LAST k-block of a
// tile's contraction needs the tile right above it.  Tiles are therefore separate tasks ordered by anti-diagonal
// (I - J), claimed in list order by persistent CTAs, and the producer warp acquires the flag of tile (K, J) just before
// it issues the k-block that reads it: the tiles of a column run concurrently on different SMs as a wavefront, and the
// critical path of an expert with nb block columns is nb x (one k-block + epilogue) instead of nb^2 / 2 k-blocks.
// (trtri2's one-CTA-per-column tasks had the long chain: 13 ms for n = 5008, which capped multi-GPU strong scaling.)
//
// The engine computes the TRANSPOSE  OUT[c][r] = S^T  (A operand = X^T rows of block J, B operand = L rows of block I)
// so that a warp owns complete rows c and the right-multiplication by W_I^T runs in registers; X_IJ^T is stored in the
// strict upper block triangle of the factor (rows of block J, columns of block I), which is also where the tiles
// further down the column read their A operand from.
// Per-tile partials (deterministic, no atomics): tpart[tile] = ||X_IJ||_F^2, apart[tile][c] = (X_IJ^T z_I)[c];
// alpha_reduce_kernel sums them per block column in a fixed order and adds W_J^T z_J.
#pragma once
#include "engine2.cuh"
#include "args.h"
#include "potrf2_args.h"

namespace dsm {

__device__ __forceinline__ int tri_tile_index(int I, int J) { return I * (I + 1) / 2 + J; }

struct Trtri3Gen {
  const double* F; const double* W; const double* WT; const int* flags;
  int nkc, I, J, nmain, nepi, c;
  bool allready;
  const int* done;            // fused evaluation kernel: flags (J, J) / (I, I) of the factorisation's diagonal tasks, else null
  TaskHdr h;
  __device__ __forceinline__ void load(const Trtri3Args& a, int ti, int kind = 0, bool wait_factor = false) {
    const int4 tk = a.tasks[ti];
    const LeafMeta m = a.meta[tk.x];
    I = tk.y; J = tk.z; c = 0; allready = false;
    F = a.F + m.foff; W = a.W + m.woff; WT = a.WT + m.woff; nkc = m.nkc;
    flags = a.flags + a.flag_off[tk.x];
    done = wait_factor ? flags : nullptr;
    h.kind = kind; h.ti = ti; h.slot = tk.x; h.I = I; h.J = J;
    h.wi = blk_width(m.np, I); h.wj = blk_width(m.np, J);
    nmain = (I - J) * (BLK / KC);
    nepi = tri_epilogue_nstages(h.wi / 32);
    h.n_c = 0; h.n_main = nmain;
  }
  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (c >= nmain + nepi) return false;
    d.flag0 = nullptr; d.flag1 = nullptr;
    const int n1 = BLK / KC;
    if (c < n1) {                                        // K = J block: X_JJ = W_J
      d.a = WT + (int64_t)J * WBLK_D + c * TILE_D; d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, J * n1 + c, nkc); d.bbytes = TILE_BYTES;
      if (c == 0 && done != nullptr) {                   // fused kernel: block row I of L, W_I, z_I and W_J^T must be final
        d.flag0 = done + tri_tile_index(J, J); d.flag1 = done + tri_tile_index(I, I);
      }
    } else if (c < nmain) {
      const int kc = J * n1 + c, K = kc / n1;
      if ((c & (n1 - 1)) == 0 && !allready) {
        if (K == J + 1 && I - J > 2) {                   // fast path: the tile right above complete => the whole column above is
          int v = 1;
          if ((threadIdx.x & 31) == 0) v = ld_acquire(flags + tri_tile_index(I - 1, J));
          allready = __shfl_sync(0xffffffffu, v, 0) != 0;
          if (allready) fence_proxy_async();
        }
        if (!allready) d.flag0 = flags + tri_tile_index(K, J);
      }
      d.a = F + tile_off(J, kc, nkc); d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, kc, nkc); d.bbytes = TILE_BYTES;
    } else {
      d = tri_epilogue_chunk(W + (int64_t)I * WBLK_D, c - nmain, nepi, nullptr);
    }
    c++;
    return true;
  }
};

__device__ __forceinline__ bool trtri3_produce_task(Pipe& p, const Trtri3Args& a, Trtri3Gen& gen, int ti, int kind, bool wait_factor) {
  if (a.mask != nullptr && a.mask[a.tasks[ti].x] == 0) return true;      // expert without a gradient request
  gen.load(a, ti, kind, wait_factor);
  ChunkDesc d;
  bool first = true;
  while (gen.next(d)) { p.issue(d, first ? &gen.h : nullptr); first = false; }
  return !*p.abort;
}

__device__ __forceinline__ void trtri3_producer(Pipe& p, const Trtri3Args& a) {
  Trtri3Gen gen;
  for (;;) {
    int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(a.counter, 1);
    const int ti = __shfl_sync(0xffffffffu, t, 0);
    if (ti >= a.ntasks) break;
    if (!trtri3_produce_task(p, a, gen, ti, 0, false)) break;
  }
  TaskHdr h; h.kind = -1;
  ChunkDesc d; d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
  p.issue(d, &h);
}

// Consumer side of one inverse tile (all 8 MMA warps): `st` = stage of the task's first chunk, `hd` its header.
__device__ __forceinline__ void trtri3_consume(Pipe& p, const Trtri3Args& a, const TaskHdr& hd, int st, double* s_red, double (*s_z)[BLK]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp_slab(), r0 = 16 * slab;
  const LeafMeta m = a.meta[hd.slot];
  const int I = hd.I, J = hd.J, i0 = I * BLK, j0 = J * BLK, wi = hd.wi;
  double* F = a.F + m.foff;
  const double* z = a.z + m.voff;
  Acc2 acc;
  acc2_zero(acc);
  {   // z_I for the fused alpha partial: latency hidden behind the contraction
    double4 zv = make_double4(0.0, 0.0, 0.0, 0.0);
    if (4 * lane < wi) {          // through L2: in the fused kernel z_I was written by another CTA of the same launch
      const double2 za = __ldcg(reinterpret_cast<const double2*>(z + i0 + 4 * lane)), zb = __ldcg(reinterpret_cast<const double2*>(z + i0 + 4 * lane + 2));
      zv = make_double4(za.x, za.y, zb.x, zb.y);
    }
    __syncwarp();
    *reinterpret_cast<double4*>(&s_z[warp][4 * lane]) = zv;
    __syncwarp();
  }
  for (int c = 0; c < hd.n_main; c++) {
    if (c > 0) st = p.wait();
    // K = J block: the A operand is W_J^T (A[row][k] = W_J[k][row] = 0 for k < row): chunk c is all zero for slabs > c
    if (c >= BLK / KC || c >= slab) {
      if (wi == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0);
    }
    p.release();
  }
  tri_epilogue(p, acc, wi / 32, true, -1.0);      // OUT = X_IJ^T  (rows c of block J, cols r of block I)
  acc2_store(acc, F, m.nkc, j0, i0, BLK, wi);
  // publish as early as possible: the tile below in this column is waiting for exactly this
  csync();
  int* flags = a.flags + a.flag_off[hd.slot];
  const int64_t tile = a.flag_off[hd.slot] + tri_tile_index(I, J);
  if (tid == 0) { __threadfence(); st_release(flags + tri_tile_index(I, J), 1); }
  // fused partials: ||X_IJ||_F^2 over real rows/cols, (X_IJ^T z_I)[c]
  double tr = 0.0, p0 = 0.0, p1 = 0.0;
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
  if ((lane & 3) == 0) *reinterpret_cast<double2*>(a.apart + tile * BLK + acc_row(0)) = make_double2(p0, p1);   // rows 2g, 2g+1 adjacent
  tr = block_sum_c(tr, s_red);
  if (tid == 0) a.tpart[tile] = tr;
}

__global__ void __launch_bounds__(NTHREADS_PW, 1) trtri3_kernel(Trtri3Args a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_red[16];
  __shared__ __align__(32) double s_z[NCONS / 32][BLK];     // per-warp copy of z_I
  const int warp = threadIdx.x >> 5;
  Pipe p;
  p.init(smem, a.gerr);
  if (warp >= NCONS / 32) {                          // producer warpgroup: one working warp, three that only donate registers
    setmaxnreg_dec<REGS_PRODUCER>();
    if (warp == NCONS / 32) trtri3_producer(p, a);
    return;
  }
  setmaxnreg_inc<REGS_CONSUMER>();
  for (;;) {
    const int st = p.wait();                         // the first chunk of a task carries its header
    const TaskHdr hd = p.hdr[st];
    if (hd.kind < 0 || *p.abort) return;
    trtri3_consume(p, a, hd, st, s_red, s_z);
  }
}

// alpha_J = W_J^T z_J + sum_{I > J} X_IJ^T z_I  and the block-column partial of tr(F^-1), in a fixed order.
// grid = block columns of the batch (the old column-task list), block = BLK threads.
__global__ void __launch_bounds__(BLK) alpha_reduce_kernel(Trtri3Args a, const int2* cols, int ncols) {
  const int2 ck = cols[blockIdx.x];
  if (a.mask != nullptr && a.mask[ck.x] == 0) return;
  const LeafMeta m = a.meta[ck.x];
  const int J = ck.y, j0 = J * BLK, wj = blk_width(m.np, J), tid = threadIdx.x;
  const double* z = a.z + m.voff;
  const double* WTj = a.WT + m.woff + (int64_t)J * WBLK_D;
  const int64_t base = a.flag_off[ck.x];
  double s = 0.0;
  if (tid < wj) for (int k = tid; k < wj; k++) s = fma(WTj[widx(tid, k)], z[j0 + k], s);
  double tr = 0.0;
  for (int I = J + 1; I < m.nb; I++) {
    const int64_t tile = base + tri_tile_index(I, J);
    s += a.apart[tile * BLK + tid];
    if (tid == 0) tr += a.tpart[tile];
  }
  if (tid < wj) a.alpha[m.voff + j0 + tid] = (j0 + tid < m.n) ? s : 0.0;
  if (tid == 0) a.trpart[a.trpart_off[ck.x] + m.nb + J] = tr;

}

}  // namespace dsm
