// Persistent, dependency-driven batched Cholesky (engine v2).
//
// Replaces LAPACK.potrf!('L', F) (gaussianprocess.jl:101) + the forward solve of gaussianprocess.jl:105 for ALL
// experts in ONE launch.  Left-looking by 128-wide block columns; a task is one 128x128 tile:
//   diag (J,J):   C = F_JJ - sum_{K<J} L_JK L_JK^T ; L_JJ = chol(C) ; W_J = L_JJ^-1 ; z_J = W_J (y_J - sum_K L_JK z_K)
//   panel (I,J):  L_IJ = (F_IJ - sum_{K<J} L_IK L_JK^T) W_J^T
// CTAs (one per SM; 8 MMA warps + 1 producer warp) claim tasks IN ORDER from a global counter.  The host writes the task list in a topological
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

// non-blocking test of a tile flag by a whole warp (lane 0 reads, result broadcast)
__device__ __forceinline__ bool flag_is_set(const int* f) {
  int v = 1;
  if ((threadIdx.x & 31) == 0) v = ld_acquire(f);
  return __shfl_sync(0xffffffffu, v, 0) != 0;
}

// Chunk generator of one tile task (producer warp).  Stage order: the tile F_IJ itself (2 column tiles per stage, no
// dependency), the contraction chunks (k-block Kb needs tiles (I,Kb) and (J,Kb)), then W_J for the TRSM epilogue
// (panel tiles only).  Dependencies are attached to the descriptor; Pipe::issue waits for them.
struct PotrfGen {
  const double* F; const double* z; const double* Wj; const int* flags;
  int nkc, I, J;
  int nc, nmain, nepi, c, cc0;
  bool diag, allready;       // allready: every k-block is known to be complete (tile (.,J-1) done implies all before it)
  bool prefd;                // diagonal tile of a block column that already holds a valid factor (chol_continue / shared prefix)
  TaskHdr h;
  __device__ __forceinline__ int total() const { return nc + nmain + nepi; }

  __device__ __forceinline__ void load(const Potrf2Args& a, int ti) {
    c = 0; nc = nmain = nepi = 0; diag = false; allready = false; prefd = false;
    const int4 tk = a.tasks[ti];
    const LeafMeta m = a.meta[tk.x];
    I = tk.y; J = tk.z; diag = (I == J);
    flags = a.flags + a.flag_off[tk.x];
    const int js = (a.share != nullptr) ? a.share[tk.x].z : a.jstart;     // block rows < js hold a valid factor
    h.kind = diag ? 1 : 0; h.ti = ti; h.slot = tk.x; h.I = I; h.J = J;
    h.wi = blk_width(m.np, I); h.wj = blk_width(m.np, J); h.n_c = 0; h.n_main = 0; h.pad0 = js; h.pad1 = 0; cc0 = 0;
    if (!diag && I < js) return;                          // tile already final (chol_continue / copied from the source expert)
    nkc = m.nkc;
    F = a.F + m.foff; z = a.z + m.voff; Wj = a.W + m.woff + (int64_t)J * WBLK_D;
    nc = h.wj / 32;
    prefd = diag && J < js;       // streams its block row for the forward solve z_J only (no MMA), then rebuilds W_J
    // right-looking split: the leading k-blocks are already in the tile.  Panel tiles skip them; a diagonal tile still
    // streams its whole block row for the forward solve z_J but runs no MMA on the first pad1 chunks
    const int ks = (a.kskip != nullptr) ? a.kskip[tk.x] : 0;
    const int skip = (ks > 0 && J >= ks && !prefd) ? ks * (BLK / KC) : 0;
    cc0 = diag ? 0 : skip; h.pad1 = diag ? skip : 0;
    nmain = (J * BLK) / KC - cc0;
    nepi = diag ? 0 : tri_epilogue_nstages(h.wj / 32);
    h.n_c = nc; h.n_main = nmain;
  }

  __device__ __forceinline__ bool next(ChunkDesc& d) {
    if (c >= nc + nmain + nepi) return false;
    d.flag0 = nullptr; d.flag1 = nullptr;
    if (c < nc) {
      d.a = F + tile_off(I, J * 8 + 2 * c, nkc); d.abytes = TILE_BYTES;
      d.b = F + tile_off(I, J * 8 + 2 * c + 1, nkc); d.bbytes = TILE_BYTES;
    } else if (c < nc + nmain) {
      const int cc = cc0 + c - nc, Kb = cc >> 3;
      if ((cc & 7) == 0 && !allready && prefd) {
        // the tiles (J, K) are final, but z_K is written by the diagonal task of column K: its flag orders all of them
        if (J > 0) d.flag0 = flags + tile_flag_index(J - 1, J - 1);
        allready = true;
      }
      if ((cc & 7) == 0 && !allready) {
        if (c == nc && J > 1) {                           // fast path: the last k-block's tiles complete => all are
          int v0 = 1, v1 = 1;                             // both loads in flight before either is consumed
          if ((threadIdx.x & 31) == 0) {
            v0 = ld_acquire(flags + tile_flag_index(I, J - 1));
            if (!diag) v1 = ld_acquire(flags + tile_flag_index(J, J - 1));
          }
          allready = __shfl_sync(0xffffffffu, (v0 != 0) && (v1 != 0) ? 1 : 0, 0) != 0;
          if (allready) fence_proxy_async();
        }
        if (!allready) {
          d.flag0 = flags + tile_flag_index(I, Kb);
          d.flag1 = diag ? nullptr : flags + tile_flag_index(J, Kb);
        }
      }
      d.a = F + tile_off(I, cc, nkc); d.abytes = TILE_BYTES;
      if (diag) { d.b = z + KC * cc; d.bbytes = KC * 8; }          // B operand == A operand; B part carries z[16cc..]
      else { d.b = F + tile_off(J, cc, nkc); d.bbytes = TILE_BYTES; }
    } else {
      const int e = c - nc - nmain;
      d = tri_epilogue_chunk(Wj, e, nepi, (e == 0) ? flags + tile_flag_index(J, J) : nullptr);
    }
    c++;
    return true;
  }
};

// Producer warp: claims tasks in list order and streams their chunks into the ring, running ahead of the MMA warps by
// up to NS2 stages ACROSS task boundaries.  A diagonal tile reuses the ring as scratch for its factorisation, so after
// the last chunk of a diagonal task the producer waits until the consumers hand the ring back (aux[0]).
// One claimed task: stream its chunks into the ring.  Returns false when the kernel must stop (abort).
__device__ __forceinline__ bool potrf2_produce_task(Pipe& p, const Potrf2Args& a, PotrfGen& gen, int ti, uint32_t& scratch_phase) {
  if (a.share != nullptr && a.share[a.tasks[ti].x].x == SHARE_ALIAS) return true;     // the source expert's results are reused
  gen.load(a, ti);
  if (gen.total() == 0) {                               // already final: publish and move on
    if ((threadIdx.x & 31) == 0) st_release(const_cast<int*>(gen.flags) + tile_flag_index(gen.I, gen.J), 1);
    return true;
  }
  ChunkDesc d;
  bool first = true;
  while (gen.next(d)) { p.issue(d, first ? &gen.h : nullptr); first = false; }
  if (gen.diag) { p.wait_bar(&p.aux[0], scratch_phase & 1, 5); scratch_phase++; }
  return !*p.abort;
}

__device__ __forceinline__ void pipe_end(Pipe& p) {       // header with kind < 0 ends the consumers
  TaskHdr h; h.kind = -1;
  ChunkDesc d; d.a = nullptr; d.b = nullptr; d.abytes = 0; d.bbytes = 0; d.flag0 = nullptr; d.flag1 = nullptr;
  p.issue(d, &h);
}

__device__ __forceinline__ void potrf2_producer(Pipe& p, const Potrf2Args& a) {
  PotrfGen gen;
  uint32_t scratch_phase = 0;
  for (;;) {
    int t = 0;
    if ((threadIdx.x & 31) == 0) t = atomicAdd(a.counter, 1);
    const int ti = __shfl_sync(0xffffffffu, t, 0);
    if (ti >= a.ntasks) break;
    if (!potrf2_produce_task(p, a, gen, ti, scratch_phase)) break;
  }
  pipe_end(p);
}

// Consumer side of one task (all 8 MMA warps): `st` = stage of the task's first chunk, `hd` its header.
// done_flags (or null): per-tile flag array in which a diagonal task marks entry (J, J) once EVERYTHING it writes (L_JJ, W_J,
// W_J^T, z_J, the partial sums) is in memory -- the inverse tiles of the fused evaluation kernel wait on it.
__device__ __forceinline__ void potrf2_consume(Pipe& p, const Potrf2Args& a, const TaskHdr& hd, int st, double* smem,
                                               double* s_red, double* s_v, int* s_info_p, int* done_flags) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slab = warp_slab(), r0 = 16 * slab;
  int& s_info = *s_info_p;
  const LeafMeta m = a.meta[hd.slot];
  const int I = hd.I, J = hd.J;
  const int i0 = I * BLK, j0 = J * BLK;
  const int wi = hd.wi, wj = hd.wj;
  double* F = a.F + m.foff;
  int* flags = a.flags + a.flag_off[hd.slot];
  double* Wj = a.W + m.woff + (int64_t)J * WBLK_D;
  const int nkc = m.nkc;
  const bool diag = (hd.kind == 1);
  const bool active = r0 < wi;
  const bool prefactored = (J < hd.pad0);      // chol_continue / shared prefix: column already final, diag only rebuilds W and z
  const int n_c = hd.n_c, n_main = hd.n_main;
  long long* trc = (a.trace != nullptr && tid == 0) ? a.trace + (long long)hd.ti * 8 : nullptr;
  if (trc) { trc[0] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); trc[5] = (long long)sm | ((long long)I << 16) | ((long long)J << 32); }

  // acc = -F_IJ: the tile arrives through the ring as the first wj/32 stages (two 16-column tiles per stage)
  Acc2 acc;
  acc2_zero(acc);
#pragma unroll
  for (int e = 0; e < 4; e++) {
    if (e < n_c) {
      if (e > 0) st = p.wait();
      if (active) { acc2_sub_tile(acc, p.A(st), r0, 2 * e); acc2_sub_tile(acc, p.B(st), r0, 2 * e + 1); }
      p.release();
    }
  }

  if (trc) trc[1] = clock64();
  if (!diag) {
    // ---------------- panel tile ----------------
    for (int c = 0; c < n_main; c++) {
      st = p.wait();
      if (active) { if (wj == BLK) mma_chunk<4>(acc, p.A(st), p.B(st), r0); else mma_chunk<2>(acc, p.A(st), p.B(st), r0); }
      p.release();
    }
    if (trc) trc[2] = clock64();
    // X = C W_J^T = (-acc) W_J^T, in registers (M = W_J streamed through the ring)
    tri_epilogue(p, acc, wj / 32, active, -1.0);
    if (trc) trc[3] = clock64();
    acc2_store(acc, F, nkc, i0, j0, wi, wj);
  } else {
  // ---------------- diagonal tile ----------------
  double gemv = 0.0;                                  // row r0 + (lane & 15), k-half lane >> 4
  {
    const int ng = min(wj / 16, slab + 1);            // lower triangle only: 16-column groups 0 .. slab
    for (int c = 0; c < n_main; c++) {
      st = p.wait();
      if (active) {
        const double* sA = p.A(st);
        if (!prefactored && c >= hd.pad1) switch (ng) {
          case 1: mma_chunk16<1>(acc, sA, sA, r0); break;
          case 2: mma_chunk16<2>(acc, sA, sA, r0); break;
          case 3: mma_chunk16<3>(acc, sA, sA, r0); break;
          case 4: mma_chunk16<4>(acc, sA, sA, r0); break;
          case 5: mma_chunk16<5>(acc, sA, sA, r0); break;
          case 6: mma_chunk16<6>(acc, sA, sA, r0); break;
          case 7: mma_chunk16<7>(acc, sA, sA, r0); break;
          default: mma_chunk16<8>(acc, sA, sA, r0); break;
        }
        const double* zs = p.B(st);
        const double* ar = sA + r0 + (lane & 15) + (lane >> 4) * 8 * LDS;
#pragma unroll
        for (int k = 0; k < 8; k++) gemv = fma(ar[k * LDS], zs[(lane >> 4) * 8 + k], gemv);
      }
      p.release();
    }
  }
  if (trc) trc[2] = clock64();
  gemv += __shfl_xor_sync(0xffffffffu, gemv, 16);
  csync();                                             // every warp is done with the ring: stages become scratch
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
  csync();
  {
    const int info = diag_factor_invert(S, wj, aux, !prefactored, &s_info, trc ? trc + 6 : nullptr);
    if (tid == 0 && info != 0) atomicCAS(&a.scal[hd.slot].info, 0, j0 + info);
  }
  if (trc) trc[3] = clock64();
  const double* DI = aux;
  // Critical path first: the panel tiles of this column wait for W_J and the next diagonal tile (transitively) for
  // z_J, so those two are stored and the tile is PUBLISHED before the factor, W_J^T and the reductions are written.
  double tr = 0.0;
  for (int c = warp; c < BLK; c += NCONS / 32)
    for (int r = lane; r < BLK; r += 32) {
      double v = 0.0;
      if (r < wj && c < wj && r >= c) v = diag_W(S, DI, r, c);
      Wj[widx(r, c)] = v;
      if (j0 + r < m.n && j0 + c < m.n) tr += v * v;
    }
  // forward solve block: z_J = W_J (y_J - sum_K L_JK z_K)
  if (tid < BLK) s_v[tid] = (tid < wj) ? a.y[m.voff + j0 + tid] - s_v[tid] : 0.0;
  csync();
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
  csync();
  if (tid == 0) { __threadfence(); st_release(flags + tile_flag_index(I, J), 1); }
  // off the critical path: L_JJ, W_J^T, log-det and the partial sums
  if (!prefactored) {
    for (int c = warp; c < wj; c += NCONS / 32) {
      double* dst = F + tidx(j0, j0 + c, nkc);
      for (int r = lane; r < wj; r += 32)
        if (r >= c) dst[r] = S[c * LDS + r];
    }
  }
  double* WTj = a.WT + m.woff + (int64_t)J * WBLK_D;
  for (int r = warp; r < BLK; r += NCONS / 32)
    for (int c = lane; c < BLK; c += 32) {
      double v = 0.0;
      if (r < wj && c < wj && r >= c) v = diag_W(S, DI, r, c);
      WTj[widx(c, r)] = v;
    }
  double ld = 0.0;
  for (int r = tid; r < wj; r += NCONS)
    if (j0 + r < m.n) ld += log(S[r * LDS + r]);
  ld = block_sum_c(ld, s_red);
  tr = block_sum_c(tr, s_red);
  zz = block_sum_c(zz, s_red);
  if (tid == 0) {
    const int64_t po = a.trpart_off[hd.slot];
    a.trpart[po + J] = tr;
    a.ldpart[po / 2 + J] = 2.0 * ld;
    a.zzpart[po / 2 + J] = zz;
  }
  fence_proxy_async();                                 // generic writes to the stages precede the next bulk copies
  }   // diagonal tile
  // ---------------- task boundary: publish a panel tile / hand the ring back to the producer after a diagonal tile
  csync();
  if (tid == 0) {
    if (!diag) { __threadfence(); st_release(flags + tile_flag_index(I, J), 1); }
    else {
      if (done_flags != nullptr) { __threadfence(); st_release(done_flags + a.flag_off[hd.slot] + tile_flag_index(J, J), 1); }
      mbar_arrive(&p.aux[0]);
    }
  }
  if (trc) trc[4] = clock64();
}

__global__ void __launch_bounds__(NTHREADS_PW, 1) potrf2_kernel(Potrf2Args a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_red[16];
  __shared__ double s_v[BLK];
  __shared__ int s_info;
  const int warp = threadIdx.x >> 5;
  Pipe p;
  p.init(smem, a.gerr);
  if (warp >= NCONS / 32) {                          // producer warpgroup: one working warp, three that only donate registers
    setmaxnreg_dec<REGS_PRODUCER>();
    if (warp == NCONS / 32) potrf2_producer(p, a);
    return;
  }
  setmaxnreg_inc<REGS_CONSUMER>();
  for (;;) {
    // the first chunk of a task carries its header
    const int st = p.wait();
    const TaskHdr hd = p.hdr[st];
    if (hd.kind < 0 || *p.abort) return;
    potrf2_consume(p, a, hd, st, smem, s_red, s_v, &s_info, nullptr);
  }
}

}  // namespace dsm
