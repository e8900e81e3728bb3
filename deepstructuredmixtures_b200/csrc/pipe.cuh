// Asynchronous operand pipeline for the FP64 DMMA engine (v2).
//
// Operand k-slices are moved global -> shared by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) and signalled
// through mbarriers, so the 8 MMA warps never meet at a block-wide barrier inside a contraction:
//   full[s]   (count 1 + tx bytes)  producer arrives with expect_tx, the copies complete the transaction
//   empty[s]  (count 8)             each warp arrives after its last shared-memory read of the stage
// A DEDICATED PRODUCER WARP (warp 8 of a 288-thread CTA) claims tasks, polls dependency flags and issues the copies; the
// eight MMA warps only wait / multiply / release, so none of the scheduler's latency (atomics, flag polls, descriptor
// loads) sits on the MMA critical path.  The first chunk of every task carries a TASK HEADER (hdr[stage]) that tells the
// consumers what the following chunks are; a header with kind < 0 ends the kernel.  Operands are stored in
// HBM as 16-column tiles that are the exact image of a stage (common.cuh), so a chunk is TWO bulk copies of 16,896 B.  Chunks are numbered monotonically per CTA
// (stage = q % NS, parity = (q / NS) & 1), so the ring keeps running across segments, epilogues and task iterations.
//
// Cross-CTA dependencies (left-looking Cholesky) are tile flags in global memory: the producer acquires the flag of
// the tile a k-block comes from before issuing its copies.  Every spin is bounded: on a timeout the CTA records an
// error code in global memory and stops waiting (results are then garbage, the host returns an error; no hang).
#pragma once
#include "common.cuh"

namespace dsm {

#ifndef DSM_EPI_EARLY
#define DSM_EPI_EARLY 1
#endif
constexpr int NS2 = 6;                                   // ring stages (6 x 33,792 B = 202,752 B)
constexpr int STAGE_DOUBLES = 2 * CHUNK;                 // A chunk then B chunk, each [KC][LDS]
constexpr int PIPE_SMEM_BYTES = NS2 * STAGE_DOUBLES * 8 + 512;   // + barriers / control words / task headers
constexpr int NTHREADS_PW = NTHREADS + 128;              // block size of the producer-warp kernels: 2 MMA warpgroups + 1 producer warpgroup
// Register re-allocation between the warpgroups (setmaxnreg works per warpgroup of 4 warps): the kernel is compiled for
// 384 threads (168 registers each); the producer group drops to 40 and the two MMA groups grow to 232.
#ifndef DSM_REGS_P
#define DSM_REGS_P 40
#endif
#ifndef DSM_REGS_C
#define DSM_REGS_C 232
#endif
constexpr int REGS_PRODUCER = DSM_REGS_P, REGS_CONSUMER = DSM_REGS_C;
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

struct TaskHdr { int kind, ti, slot, I, J, wi, wj, n_c, n_main, pad0, pad1, pad2; };   // 48 bytes
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000LL;  // ~2 s at 1.97 GHz

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One pipeline chunk = up to two bulk copies: a tile (or any 16-byte multiple) into the A part and one into the B part.
struct ChunkDesc {
  const double* a; uint32_t abytes;
  const double* b; uint32_t bbytes;
  const int* flag0; const int* flag1;    // tiles that must be complete before the copies are issued (or null)
};

struct Pipe {
  double* base;          // NS2 stages
  uint64_t* full;        // [NS2]
  uint64_t* empty;       // [NS2]
  uint64_t* aux;         // [2] auxiliary consumer -> producer barriers (scratch free / block stored)
  volatile int* abort;   // shared: set on timeout
  TaskHdr* hdr;          // [NS2] task header of the chunk in each stage (valid for the first chunk of a task)
  int* gerr;             // global error word
  uint32_t q_issue, q_cons;

  __device__ __forceinline__ double* A(int st) const { return base + st * STAGE_DOUBLES; }
  __device__ __forceinline__ double* B(int st) const { return base + st * STAGE_DOUBLES + CHUNK; }

  // Called by all threads once per kernel.  `smem` = dynamic shared memory base (16-byte aligned).
  __device__ __forceinline__ void init(double* smem, int* global_err) {
    base = smem;
    full = reinterpret_cast<uint64_t*>(smem + NS2 * STAGE_DOUBLES);
    empty = full + NS2;
    aux = empty + NS2;
    abort = reinterpret_cast<volatile int*>(aux + 2);
    hdr = reinterpret_cast<TaskHdr*>(aux + 4);
    gerr = global_err;
    q_issue = 0; q_cons = 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < NS2; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS / 32); }
      mbar_init(&aux[0], 1); mbar_init(&aux[1], 1);
      *abort = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncthreads();
  }

  __device__ __forceinline__ void fail(int code) {
    *abort = 1;
    atomicCAS(gerr, 0, code);
  }

  // bounded spin on an mbarrier phase
  __device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
      if (*abort) return;
      if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) { fail(code); return; }
    }
  }

  // bounded spin on a cross-CTA tile flag (lane 0 polls, result broadcast)
  __device__ __forceinline__ void wait_flag(const int* flag) {
    if (flag == nullptr) return;
    if ((threadIdx.x & 31) == 0) {
      if (ld_acquire(flag) == 0) {
        const long long t0 = clock64();
        while (ld_acquire(flag) == 0) {
          if (*abort) break;
          if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) { fail(2); break; }
          __nanosleep(64);
        }
      }
    }
    __syncwarp();
  }

  // Producer side (warp 0): issue one chunk into the ring.  UBLKCP is a warp-uniform instruction: lane 0 issues.
  __device__ __forceinline__ void issue(const ChunkDesc& d, const TaskHdr* h = nullptr) {
    const uint32_t q = q_issue;
    const int st = q % NS2;
    if (q >= NS2) wait_bar(&empty[st], ((q / NS2) - 1) & 1, 3);
    if (d.flag0 != nullptr || d.flag1 != nullptr) {
      wait_flag(d.flag0);
      wait_flag(d.flag1);
      fence_proxy_async();
    }
    if ((threadIdx.x & 31) == 0) {
      if (h != nullptr) hdr[st] = *h;               // ordered before the consumers' reads by the barrier's release/acquire
      mbar_expect_tx(&full[st], d.abytes + d.bbytes);
      if (d.abytes) bulk_g2s(A(st), d.a, d.abytes, &full[st]);
      if (d.bbytes) bulk_g2s(B(st), d.b, d.bbytes, &full[st]);
    }
    __syncwarp();
    q_issue = q + 1;
  }

  // Consumer side (every warp): wait for the next chunk, return its stage.
  __device__ __forceinline__ int wait() {
    const uint32_t q = q_cons;
    const int st = q % NS2;
    wait_bar(&full[st], (q / NS2) & 1, 4);
    return st;
  }
  // Consumer side: this warp has finished reading the stage of chunk q_cons.
  __device__ __forceinline__ void release() {
    const int st = q_cons % NS2;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);
    q_cons++;
  }
  __device__ __forceinline__ bool can_issue() const { return q_issue - q_cons < (uint32_t)NS2; }
};

}  // namespace dsm
