// The Cholesky and the triangular inverse of a gradient evaluation as ONE persistent launch.
//
// potrf2_kernel and trtri3_kernel are both dependency-driven task pipelines on the same engine; launched back to back, the SMs
// that run out of factorisation tiles idle until the last block column of the largest expert is done, and the inverse then
// starts with its own ramp-up.  On a large batch that is < 1 % of the time; on a multi-GPU shard (18 experts per GPU at 8 GPUs,
// where the chain of the largest expert is most of the phase) it is where the strong-scaling efficiency goes.  Here the task
// list is [factorisation tiles | inverse tiles], claimed in order from ONE counter: a CTA that finds no factorisation tile left
// takes inverse tiles of experts / block rows that are already complete.
//   * an inverse tile (I, J) reads block row I of L, W_I, z_I and W_J^T: the diagonal tasks of the factorisation mark a second
//     flag (entry (J, J) of the INVERSE's flag array) once everything they write is in memory; the producer acquires
//     (J, J) and (I, I) before it issues the tile's first chunk.  Factorisation tasks never wait on inverse tasks and precede
//     them in the list, so the "a task only waits on earlier tasks" argument still rules out deadlock.
//   * inverse tiles are ordered by the level at which their block row completes in the factorisation's order, then row-major
//     (a tile only depends on the tiles above it in its own column).
#pragma once
#include "potrf2.cuh"
#include "trtri3.cuh"

namespace dsm {

struct Eval2Args {
  Potrf2Args p;        // p.counter is the one task counter; p.tasks / p.ntasks the factorisation tiles
  Trtri3Args t;        // t.tasks / t.ntasks the inverse tiles (claimed as ti - p.ntasks); t.flags is a separate flag array
};

__global__ void __launch_bounds__(NTHREADS_PW, 1) eval2_kernel(Eval2Args a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_red[16];
  __shared__ double s_v[BLK];
  __shared__ int s_info;
  __shared__ __align__(32) double s_z[NCONS / 32][BLK];
  const int warp = threadIdx.x >> 5;
  Pipe p;
  p.init(smem, a.p.gerr);
  if (warp >= NCONS / 32) {
    setmaxnreg_dec<REGS_PRODUCER>();
    if (warp == NCONS / 32) {
      PotrfGen pgen; Trtri3Gen tgen;
      uint32_t scratch_phase = 0;
      const int ntot = a.p.ntasks + a.t.ntasks;
      for (;;) {
        int t = 0;
        if ((threadIdx.x & 31) == 0) t = atomicAdd(a.p.counter, 1);
        const int ti = __shfl_sync(0xffffffffu, t, 0);
        if (ti >= ntot) break;
        const bool ok = ti < a.p.ntasks ? potrf2_produce_task(p, a.p, pgen, ti, scratch_phase)
                                        : trtri3_produce_task(p, a.t, tgen, ti - a.p.ntasks, 2, true);
        if (!ok) break;
      }
      pipe_end(p);
    }
    return;
  }
  setmaxnreg_inc<REGS_CONSUMER>();
  for (;;) {
    const int st = p.wait();
    const TaskHdr hd = p.hdr[st];
    if (hd.kind < 0 || *p.abort) return;
    if (hd.kind == 2) trtri3_consume(p, a.t, hd, st, s_red, s_z);
    else potrf2_consume(p, a.p, hd, st, smem, s_red, s_v, &s_info, a.t.flags);
  }
}

}  // namespace dsm
