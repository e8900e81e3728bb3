// FP64-accurate block products on the INT8 tcgen05 tensor cores (sm_100a), used by the split triangular inverse
// (ozaki_args.h, api_ozaki.cu).  Stand-alone prototype with the accuracy and rate measurements: tools/ozaki_proto.cu.
//
// Ozaki split, error-free: each operand row gets one power-of-two scale 2^e (its largest magnitude over K), the scaled
// entries are cut into S signed 7-bit slices  x = 2^(e-6) * sum_s q_s 128^-s, |q_s| <= 64  (exact in FP64).  A product
// of two slices accumulates EXACTLY in an int32 TMEM accumulator (|sum| <= 4096 * 8 * K < 2^31 for K <= 65,536); the
// pairs with s + t = g share the weight 128^-g and one accumulator; pairs with s + t >= S are dropped (< 2^-7S of
// rowmax * colmax per term).  S = 8 gives products that are MORE accurate than an FP64 FMA chain (measured: 2e-15 of
// max|C| on Cholesky-factor operands against 4e-15 for cuBLAS DGEMM).
//
// GEMM kernel: persistent CTAs, one 128 x 128 output block at a time.  warp 0 = TMA producer (cp.async.bulk.tensor.2d of 4 KB slice
// tiles into a mbarrier ring), warp 1 = TMEM allocation + tcgen05.mma.kind::i8 issue (M = N = 128, K = 32), warps 2-5 =
// 8 epilogue warps (tcgen05.ld, int32 -> FP64, Horner sum over the slice groups, row scales, store in the factor-tile layout).
// S groups of 128 columns do not fit TMEM (512 columns), so a block is computed in two rounds over K: the four
// lowest-weight groups first (their partial sums stay in the epilogue warps' registers), then the remaining S - 4 groups.
#pragma once
#include <cuda.h>
#include "ozaki_args.h"

namespace dsm {
namespace oz {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol error must not hang the device; the trap surfaces as a CUDA error of the call
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (UMMA SmemDescriptor, version 1): [0,14) address >> 4,
// [16,30) byte offset between the two 16-byte K chunks >> 4, [32,46) byte offset between 8-row groups >> 4, [46,48) = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((BLK * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

template <int S>
struct Cfg {
  static constexpr int NR = (S > 4) ? 2 : 1;
  static constexpr int STAGE = 2 * S * OZ_TILE_B;
  static constexpr int NST = (220 * 1024) / STAGE;
  static constexpr int SMEM = NST * STAGE + 1024;
  __host__ __device__ static constexpr int glo(int r) { return (NR == 2 && r == 0) ? S - 4 : 0; }
  __host__ __device__ static constexpr int ghi(int r) { return (NR == 2 && r == 1) ? S - 5 : S - 1; }
  __host__ __device__ static constexpr int nsl(int r) { return ghi(r) + 1; }
};

// ---- slicing ---------------------------------------------------------------------------------------------------------
// grid = jobs, 256 threads: thread = (operand row i, 16-wide k chunk).  pass 0: row maxima (atomicMax on the bit pattern
// of |x|, monotone for non-negative doubles); pass 1: the slices, 16 bytes per thread and slice.
template <int S>
__global__ void __launch_bounds__(256) slice_kernel(const OzJob* __restrict__ jobs, int pass, unsigned long long* __restrict__ rowmax,
                                                    double* __restrict__ scale, int8_t* __restrict__ pool) {
  const OzJob j = jobs[blockIdx.x];
  const int i = threadIdx.x & 127;
  double inv = 0.0;
  if (pass == 1) {
    const double m = __longlong_as_double((long long)rowmax[j.scale + i]);
    int e = 0;
    if (m > 0.0 && m < 1.0e300) frexp(m, &e);
    inv = ldexp(64.0, -e);
    scale[j.scale + i] = ldexp(1.0, e - 6);        // every block of the row writes the same value
  }
  double mx = 0.0;
#pragma unroll 1
  for (int kc = threadIdx.x >> 7; kc < BLK / 16; kc += 2) {
    double v[16];
    if (i < j.vr) {
      if (!j.transposed) {
        const double* p = j.src + kc * TILE_D + i;
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = (16 * kc + q < j.vk) ? p[q * LDS] : 0.0;
      } else {
        const double* p = j.src + (i >> 4) * TILE_D + (i & 15) * LDS + 16 * kc;
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = (16 * kc + q < j.vk) ? p[q] : 0.0;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] = 0.0;
    }
    if (pass == 0) {
#pragma unroll
      for (int q = 0; q < 16; q++) mx = fmax(mx, fabs(v[q]));
    } else {
      int8_t* dst = pool + (j.dst + (int64_t)(kc >> 1) * S) * OZ_TILE_B + (kc & 1) * (BLK * 16) + (i >> 3) * 128 + (i & 7) * 16;
#pragma unroll
      for (int q = 0; q < 16; q++) v[q] *= inv;
#pragma unroll
      for (int s = 0; s < S; s++) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 16; q++) {
          // rint and double -> int in one FP64 add: |v| <= 64.5, so v + 1.5 * 2^52 holds rint(v) (nearest-even, like rint) in the
          // low mantissa bits as a two's-complement integer.  The conversion instructions (F2F.F64 round, F2I) run at a quarter
          // of the FP64 add rate and made this kernel conversion-bound (sm throughput 85 % at 4.1 TB/s).
          const double t = v[q] + 6755399441055744.0;
          const double r = t - 6755399441055744.0;
          v[q] = (v[q] - r) * 128.0;
          w[q >> 2] |= ((uint32_t)__double2loint(t) & 0xFFu) << ((q & 3) * 8);
        }
        *reinterpret_cast<uint4*>(dst + (int64_t)s * OZ_TILE_B) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  if (pass == 0 && mx > 0.0) atomicMax(rowmax + j.scale + i, (unsigned long long)__double_as_longlong(mx));
}

// ---- block products ----------------------------------------------------------------------------------------------------
constexpr int OZ_THREADS = 64 + 256;      // producer warp, MMA warp, 8 epilogue warps

// Persistent: CTA b works on the blocks b, b + gridDim.x, ... of the list (sorted by length, so the strided assignment is
// balanced to within one block).  TMEM is allocated once; the operand ring and the barriers run on across blocks.  The
// epilogue warps release TMEM as soon as they have read it, so the global-memory part of a block's epilogue (scales, old
// values of an accumulating block, stores) overlaps the MMAs of the next block.
template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const OzTile* __restrict__ tiles, int ntiles,
            const double* __restrict__ scale, long long* __restrict__ trace) {
  using C = Cfg<S>;
  constexpr int RING = C::NST * C::STAGE;                     // bytes of the operand ring
  constexpr int MAXST = 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  // per round its own ring geometry (a stage = the slices that round needs for one k-step) and its own barriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);         // [round][full MAXST | empty MAXST]
  uint64_t* tfull = bars + 4 * MAXST;
  uint64_t* tempty = tfull + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(tempty + 1);
  uint8_t* ring = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * MAXST; ++i) mbar_init(&bars[i], 1);
    mbar_init(tfull, 1); mbar_init(tempty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;

  if (warp == 0) {
    if (lane == 0) {
      int it[2] = {0, 0};          // steps issued into the ring of each round so far
      int nround = 0;              // (block, round) pairs started so far
      for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const OzTile t = tiles[ti];
        const int nk = t.k1 - t.k0;
#pragma unroll
        for (int r = 0; r < C::NR; ++r, ++nround) {
          const int nsl = C::nsl(r);
          const int stage = 2 * nsl * OZ_TILE_B;
          const int nst = (RING / stage) < MAXST ? (RING / stage) : MAXST;
          uint64_t* full = bars + r * 2 * MAXST; uint64_t* empty = full + MAXST;
          // the ring changes its geometry between rounds: it is free once the previous round's last MMAs have completed
          // (tfull); the loads of this round then overlap the epilogue of the previous one
          if (C::NR > 1 && nround > 0) mbar_wait(tfull, (nround - 1) & 1);
          for (int i = 0; i < nk; ++i, ++it[r]) {
            const int st = it[r] % nst, ks = t.k0 + i;
            if (it[r] >= nst) mbar_wait(&empty[st], ((it[r] / nst) - 1) & 1);
            uint8_t* dst = ring + st * stage;
            mbar_expect_tx(&full[st], stage);
            // tensor-map rows = 128-byte core matrices; the box of round r's map covers the nsl slice tiles of a k-step at once
            const int rowA = (t.a_tile + ks * S) * 32, rowB = (t.b_tile + ks * S) * 32;
            const CUtensorMap* map = (r == 0) ? &map0 : &map1;
            tma_load_2d(dst, map, 0, rowA, &full[st]);
            tma_load_2d(dst + nsl * OZ_TILE_B, map, 0, rowB, &full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), K-major, N >> 3 at 17, M >> 4 at 24
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLK >> 3) << 17) | ((uint32_t)(BLK >> 4) << 24);
      int it[2] = {0, 0};
      int nround = 0;
      for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const int nk = tiles[ti].k1 - tiles[ti].k0;
        long long* trc = trace ? trace + (int64_t)ti * 16 : nullptr;
#pragma unroll
        for (int r = 0; r < C::NR; ++r, ++nround) {
          const int glo = C::glo(r), ghi = C::ghi(r), nsl = C::nsl(r);
          const int stage = 2 * nsl * OZ_TILE_B;
          const int nst = (RING / stage) < MAXST ? (RING / stage) : MAXST;
          uint64_t* full = bars + r * 2 * MAXST; uint64_t* empty = full + MAXST;
          if (nround > 0) { mbar_wait(tempty, (nround - 1) & 1); tc_fence_after(); }     // the epilogue has read the accumulators
          if (trc && r == 0) trc[0] = clock64();
          for (int i = 0; i < nk; ++i, ++it[r]) {
            const int st = it[r] % nst;
            mbar_wait(&full[st], (it[r] / nst) & 1);
            if (trc && r == 0 && i == 0) trc[2] = clock64();
            tc_fence_after();
            const uint32_t sA = smem_u32(ring + st * stage), sB = sA + nsl * OZ_TILE_B;
            uint32_t written = (i > 0) ? 0xFFu : 0u;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              if (s >= nsl) continue;
              const uint64_t da = make_desc(sA + s * OZ_TILE_B);
#pragma unroll
              for (int u = 0; u < S; ++u) {
                const int g = s + u;
                if (u >= nsl || g < glo || g > ghi) continue;
                tc_mma_i8(tbase + (uint32_t)((g - glo) * BLK), da, make_desc(sB + u * OZ_TILE_B), idesc, (written >> g) & 1u);
                written |= 1u << g;
              }
            }
            tc_commit(&empty[st]);
          }
          tc_commit(tfull);
          if (trc) trc[3 + 2 * r] = clock64();
        }
      }
    }
  } else {
    // epilogue: 8 warps.  Warp w may touch TMEM lanes 32 (w % 4) .. +31 = rows of the block; the two warps of a lane quarter
    // split the 128 columns.  The Horner sum over the slice groups stays in registers across the two rounds.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane, cbase = half * 64;
    int nround = 0;
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
      const OzTile t = tiles[ti];
      long long* trc = (trace && threadIdx.x == 64) ? trace + (int64_t)ti * 16 : nullptr;
      const bool rok = row < t.vr;
      const double srow = rok ? scale[t.sa + row] * t.sign : 0.0;
      double acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.0;
#pragma unroll
      for (int r = 0; r < C::NR; ++r, ++nround) {
        const int glo = C::glo(r), ghi = C::ghi(r);
        mbar_wait(tfull, nround & 1);
        if (trc) trc[4 + 2 * r] = clock64();
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
#pragma unroll
          for (int g = S - 1; g >= 0; --g) {                       // smallest weight first: acc = acc / 128 + G_g
            if (g < glo || g > ghi) continue;
            uint32_t v[16];
            tc_ld16(tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - glo) * BLK + cbase + c0), v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              // int32 -> FP64 without I2F (quarter-rate conversion unit): the bits of 2^52 + (x + 2^31), minus that offset -- exact
              const double gx = __hiloint2double(0x43300000, (int)(v[j] ^ 0x80000000u)) - 4503601774854144.0;
              acc[c0 + j] = fma(acc[c0 + j], 0.0078125, gx);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tempty)) : "memory");   // TMEM is free again
      }
      if (trc) trc[7] = clock64();
      // stores in chunks of 8 columns: all loads of a chunk (column scales, old values of an accumulating block) are issued
      // before its stores -- interleaved they would serialise on the memory latency
#pragma unroll
      for (int j0 = 0; j0 < 64; j0 += 8) {
        const int c0 = cbase + j0;
        if (rok && c0 < t.vc) {                                      // vc is a multiple of 16
          double* o = t.out + (c0 >> 4) * TILE_D + (c0 & 15) * LDS + row;      // column c0 + j at o[j * LDS]
          double sc[8], old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) sc[j] = scale[t.sb + c0 + j];
#pragma unroll
          for (int j = 0; j < 8; ++j) old[j] = t.accum ? o[j * LDS] : 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j * LDS] = fma(acc[j0 + j] * srow, sc[j], old[j]);
        }
      }
      if (trc) trc[8] = clock64();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
  }
}

// ---- fused partials of the inverse tiles written by the GEMMs (what trtri3_consume computes in its epilogue) ------------
// tpart[tile] = ||X_IJ||_F^2 over real rows / columns, apart[tile][c] = (X_IJ^T z_I)[c].  The block holds X_IJ^T:
// (r = column c of X_IJ, col = row i of X_IJ).  One CTA of BLK threads per tile, fixed summation order.
__global__ void __launch_bounds__(BLK) parts_kernel(OzPartArgs a) {
  __shared__ double s_red[BLK / 32];
  const OzPart p = a.parts[blockIdx.x];
  const LeafMeta m = a.meta[p.slot];
  const int i0 = p.I * BLK, j0 = p.J * BLK, wi = blk_width(m.np, p.I), tid = threadIdx.x;
  const double* blk = a.F + m.foff + tile_off(p.J, p.I * (BLK / KC), m.nkc);
  const double* z = a.z + m.voff + i0;
  const int64_t tile = a.flag_off[p.slot] + (int64_t)p.I * (p.I + 1) / 2 + p.J;
  double tr = 0.0, s = 0.0;
  const bool rowok = j0 + tid < m.n;
  for (int i = 0; i < wi; i++) {
    if (i0 + i < m.n && rowok) {
      const double v = blk[(i >> 4) * TILE_D + (i & 15) * LDS + tid];
      tr = fma(v, v, tr); s = fma(v, z[i], s);
    }
  }
  a.apart[tile * BLK + tid] = s;
  tr = warp_sum(tr);
  if ((tid & 31) == 0) s_red[tid >> 5] = tr;
  __syncthreads();
  if (tid == 0) a.tpart[tile] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
}

// Marks the factor tiles (I, J) of a part list as complete in a tile-flag array (the L21 tiles written by a GEMM instead of
// panel tasks: the diagonal tasks of the second factorisation launch stream them for the forward solve).
__global__ void setflags_kernel(const OzPart* __restrict__ parts, int n, const int64_t* __restrict__ flag_off, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const OzPart p = parts[i];
  flags[flag_off[p.slot] + (int64_t)p.I * (p.I + 1) / 2 + p.J] = 1;
}

}  // namespace oz
}  // namespace dsm
