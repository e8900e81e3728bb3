// Shared definitions for libdsmgp (sm_100a only; no other architecture is built or supported).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace dsm {

// ---- blocking constants ---------------------------------------------------------------------
// Every expert matrix is padded to a multiple of PAD (identity on the padded diagonal) and is
// processed in BLK x BLK macro tiles; the last macro tile of a matrix may be half wide (64).
constexpr int PAD = 64;
constexpr int BLK = 128;
constexpr int KC = 16;             // k-chunk staged per pipeline stage
constexpr int LDS = 132;           // smem row stride in doubles (132 % 16 == 4 -> conflict-free DMMA fragment loads)
constexpr int CHUNK = KC * LDS;    // doubles per operand per stage
constexpr int NTHREADS = 256;      // MMA (consumer) threads of the engine kernels; block size of the small kernels

enum KernelType : int { ISO_SE = 0, ARD_SE = 1, ISO_LINEAR = 2, ARD_LINEAR = 3 };

// ---- chunk-tiled matrix layout -----------------------------------------------------------------
// Factors live in HBM as a grid of TILES: tile (rb, kc) = rows [128 rb, 128 rb + 128) x columns [16 kc, 16 kc + 16),
// stored k-major with the shared-memory row stride:  tile[kk * LDS + r].  A tile is therefore the exact image of one
// operand chunk of a pipeline stage and moves global -> shared with ONE bulk copy (cp.async.bulk, 16,896 B).
// Tiles of one row block are contiguous in kc (the streaming direction of every contraction).
constexpr int TILE_D = KC * LDS;                 // doubles per tile (== CHUNK)
constexpr int TILE_BYTES = TILE_D * 8;
constexpr int WBLK_D = (BLK / KC) * TILE_D;      // a 128 x 128 diagonal-block inverse: 8 tiles
__host__ __device__ __forceinline__ int64_t tiled_doubles(int np) { return (int64_t)((np + BLK - 1) / BLK) * (np / KC) * TILE_D; }
__host__ __device__ __forceinline__ int64_t tidx(int r, int c, int nkc) {
  return ((int64_t)(r >> 7) * nkc + (c >> 4)) * TILE_D + (c & 15) * LDS + (r & 127);
}
__host__ __device__ __forceinline__ int64_t tile_off(int rb, int kc, int nkc) { return ((int64_t)rb * nkc + kc) * TILE_D; }
__host__ __device__ __forceinline__ int widx(int r, int k) { return (k >> 4) * TILE_D + (k & 15) * LDS + r; }   // inside a W block

// Per-leaf metadata, device resident.  Offsets are in doubles.
struct LeafMeta {
  int32_t n;        // expert size
  int32_t np;       // padded size (multiple of PAD) == leading dimension of the factor
  int32_t nb;       // number of macro blocks = ceil(np / BLK)
  int32_t kid;      // kernel id
  int32_t ktype;    // KernelType
  int32_t leaf;     // global leaf number
  int32_t nl;       // number of length-scale parameters (1 or D)
  int32_t nkc;      // np / 16: column chunks per row block of the tiled factor
  int64_t foff;     // factor arena offset (tiled layout, tiled_doubles(np) doubles)
  int64_t voff;     // y / z / alpha offset (length np, zero padded)
  int64_t xoff;     // gathered inputs: D columns of length np
  int64_t woff;     // inverse-diagonal-block buffers W / WT: nb blocks of WBLK_D doubles (tiled)
  int64_t poff;     // derived parameter block
};

// Derived per-leaf parameters (written by the host in set_params), layout at prm + poff:
//   [0..nl)      coef_d : SE kernels  -0.5 / l_d^2 ;  linear kernels 1 / l_d^2
//   [PRM_V]      v      = exp(2 log sigma)   (1 for linear kernels)
//   [PRM_S]      s      = exp(log sigma)     (1 for linear kernels)
//   [PRM_ETA]    eta    = exp(2 logNoise)
//   [PRM_C]      c      = eta + 1e-8
constexpr int PRM_V = 0, PRM_S = 1, PRM_ETA = 2, PRM_C = 3, PRM_COEF = 4;

// Per-leaf scalar results, device resident (one struct per local leaf)
struct LeafScal {
  double logdet;    // 2 * sum log L_ii
  double zz;        // z'z = y' alpha
  double aa;        // alpha' alpha
  double trinv;     // tr(F^{-1})
  int32_t info;     // LAPACK potrf info (0 ok, k>0 first non-positive pivot)
  int32_t pad_;
};

__device__ __forceinline__ int blk_width(int np, int b) { int w = np - b * BLK; return w < BLK ? w : BLK; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (NTHREADS threads); result valid in every thread.  `red` >= 8 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < NTHREADS / 32; i++) t += red[i];
  __syncthreads();
  return t;
}

constexpr int NCONS = NTHREADS;     // consumer (MMA) threads of the producer-warp kernels; the producer warp follows them

// block barrier of the 8 consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Block-wide sum over the consumer threads; result valid in every consumer thread.  `red` >= 8 doubles of smem.
__device__ __forceinline__ double block_sum_c(double v, double* red) {
  v = warp_sum(v);
  csync();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  csync();
  double t = 0;
#pragma unroll
  for (int i = 0; i < NCONS / 32; i++) t += red[i];
  csync();
  return t;
}

}  // namespace dsm
