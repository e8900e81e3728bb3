// tools/ozaki_proto.cu — go / no-go prototype for DESIGN.md §8 item 0.
//
// Question: can the GEMM-shaped part of the per-expert factorisation (the trailing update  C -= L_I · L_Jᵀ, 95 % of
// the flops of potrf2 / trtri3 / lauum3) run on the INT8 tcgen05 tensor cores of sm_100a with FP64-equivalent results,
// and how fast compared with the DMMA path (cuBLAS DGEMM 35.9 TFLOP/s on this pool)?
//
// Scheme (Ozaki splitting, error-free):  every row of an operand gets one power-of-two scale 2^e (its largest
// magnitude), the scaled entries v in (-1, 1) are split into S signed 7-bit slices  v = sum_s q_s / (64 * 128^s),
// |q_s| <= 64 (exact in FP64: scalings by powers of two, rint, subtraction).  A product of two slices accumulates
// EXACTLY in the int32 TMEM accumulator (|sum| <= 4096 * pairs * K < 2^31 for K <= 65,536); the slice pairs with
// s + t = g share the weight 128^-g and one accumulator, pairs with s + t >= S are dropped (below 2^-(7S) of
// rowmax * colmax).  The epilogue reads the S accumulators with tcgen05.ld, converts int32 -> FP64 and sums them
// smallest weight first, then applies the two row scales.
//
// Kernel: one CTA per 128 x 64 tile of C; warp 0 = producer (cp.async.bulk of pre-tiled slices, mbarrier ring),
// warp 1 = TMEM allocation + tcgen05.mma issue (kind::i8, M = 128, N = 64, K = 32 per instruction, S(S+1)/2
// instructions per K-step of 32), warps 2-5 = epilogue.  Operand tiles are stored by the slicing kernel in the
// UMMA canonical K-major no-swizzle layout (8 x 16-byte core matrices), so ONE bulk copy per operand and K-step
// group lands all slices in shared memory and no tensor map / swizzle is involved.
//
// Build (GPU box or here):  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/ozaki_proto tools/ozaki_proto.cu -lcublas
// Run:  /tmp/ozaki_proto [n] [K] [factor.bin]     (tools/ozaki_run.sh drives it on the GPU box)
#include <cublas_v2.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int TM = 128;        // rows of A per tile (TMEM lanes)
constexpr int TN = 128;        // rows of B per tile (accumulator columns per slice group)
constexpr int KSTEP = 32;      // K of one tcgen05.mma kind::i8
constexpr int A_TILE = TM * KSTEP;  // 4096 bytes: [2 k-chunks of 16 B][16 row groups][8 rows][16 B]
constexpr int B_TILE = TN * KSTEP;  // 4096 bytes, same layout
constexpr int SMEM_BUDGET = 220 * 1024;

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
// bounded wait: a protocol error traps instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("ozaki: barrier timeout tag %d block (%d,%d)\n", tag, blockIdx.x, blockIdx.y); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout, version 1):
// [0,14) address >> 4, [16,30) leading byte offset >> 4 (between the two 16-byte K chunks),
// [32,46) stride byte offset >> 4 (between 8-row groups), [46,48) version = 1, [61,64) layout type 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// ------------------------------------------------------------------------------------------------------------------
// Slicing: X is R x K row-major (K contiguous).  One warp per row; a lane converts 16 consecutive K entries per step
// and stores one 16-byte vector per slice.  Output tile order: [row block][K step][slice][tile bytes].
// scale[r] = 2^e / 64 so that  x = scale * sum_s q_s * 128^-s.
template <int S, int TR>
__global__ void slice_kernel(const double* __restrict__ X, int R, int K, long long ldx, int8_t* __restrict__ out, double* __restrict__ scale) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const double* x = X + (long long)row * ldx;
  double m = 0.0;
  for (int k = lane; k < K; k += 32) m = fmax(m, fabs(x[k]));
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  int e = 0;
  if (m > 0.0) frexp(m, &e);                       // m = f * 2^e, f in [0.5, 1)  ->  |x| * 2^-e < 1
  double inv = ldexp(64.0, -e);                    // first slice: q0 = rint(x * 2^-e * 64)
  if (lane == 0) scale[row] = ldexp(1.0, e - 6);
  int nk = K / KSTEP, rb = row / TR, rr = row % TR;
  constexpr int TILE = TR * KSTEP;
  for (int c = lane; c < K / 16; c += 32) {
    int ks = c >> 1, kc = c & 1;
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = x[c * 16 + j] * inv;
    int8_t* dst = out + ((long long)(rb * (long long)nk + ks) * S) * TILE + kc * (TR * 16) + (rr >> 3) * 128 + (rr & 7) * 16;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double q = rint(v[j]);
        v[j] = (v[j] - q) * 128.0;
        w[j >> 2] |= ((uint32_t)(uint8_t)(int8_t)(int)q) << ((j & 3) * 8);
      }
      *reinterpret_cast<uint4*>(dst + (long long)s * TILE) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <int S>
struct GemmCfg {
  static constexpr int NR = (S > 4) ? 2 : 1;                     // rounds: TMEM holds 4 accumulators of 128 columns
  static constexpr int STAGE = S * (A_TILE + B_TILE);
  static constexpr int NST = SMEM_BUDGET / STAGE;
  static constexpr int SMEM = NST * STAGE + 1024;
  static constexpr int TCOLS = 512;
  // round r: groups [glo, ghi], slices 0 .. nsl-1.  Low-weight groups first.
  __host__ __device__ static constexpr int glo(int r) { return (NR == 2 && r == 0) ? S - 4 : 0; }
  __host__ __device__ static constexpr int ghi(int r) { return (NR == 2 && r == 1) ? S - 5 : S - 1; }
  __host__ __device__ static constexpr int nsl(int r) { return ghi(r) + 1; }
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// One CTA per 128 x 128 tile of C.  The S slice groups do not fit TMEM at N = 128 (S * 128 columns > 512), so the tile
// is computed in two rounds over K: first the four lowest-weight groups (all slices), then the remaining S - 4 groups
// (slices 0 .. S-5); the partial sum of the first round is parked in C and folded into the Horner sum of the second.
template <int S>
__global__ void __launch_bounds__(192, 1)
ozaki_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const double* __restrict__ sa, const double* __restrict__ sb, double* __restrict__ C, long long ldc, int nk) {
  using Cfg = GemmCfg<S>;
  constexpr int NST = Cfg::NST;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(tempty + 1);
  uint8_t* stage0 = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = blockIdx.x, rb = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"((uint32_t)Cfg::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tptr;

  if (warp == 0) {
    if (lane == 0) {
      int step = 0;
      for (int r = 0; r < Cfg::NR; ++r) {
        const int nsl = Cfg::nsl(r);
        for (int ks = 0; ks < nk; ++ks, ++step) {
          int st = step % NST;
          if (step >= NST) mbar_wait(&empty[st], ((step / NST) - 1) & 1, 1);
          uint8_t* dst = stage0 + st * Cfg::STAGE;
          mbar_expect_tx(&full[st], nsl * (A_TILE + B_TILE));
          // tensor-map rows are 128-byte core matrices; one slice tile = 32 rows
          int rowA = ((rb * nk + ks) * S) * 32, rowB = ((cb * nk + ks) * S) * 32;
          for (int s = 0; s < nsl; ++s) {
            tma_load_2d(dst + s * A_TILE, &mapA, 0, rowA + s * 32, &full[st]);
            tma_load_2d(dst + S * A_TILE + s * B_TILE, &mapB, 0, rowB + s * 32, &full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      int step = 0;
#pragma unroll
      for (int r = 0; r < Cfg::NR; ++r) {
        constexpr int dummy = 0; (void)dummy;
        const int glo = Cfg::glo(r), ghi = Cfg::ghi(r), nsl = Cfg::nsl(r);
        if (r > 0) { mbar_wait(tempty, (r - 1) & 1, 4); tc_fence_after(); }
        for (int ks = 0; ks < nk; ++ks, ++step) {
          int st = step % NST;
          mbar_wait(&full[st], (step / NST) & 1, 2);
          tc_fence_after();
          uint32_t sA = smem_u32(stage0 + st * Cfg::STAGE), sB = sA + S * A_TILE;
          uint32_t written = (ks > 0) ? 0xFFu : 0u;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if (s >= nsl) continue;
            uint64_t da = make_desc(sA + s * A_TILE, TM * 16, 128);
#pragma unroll
            for (int t = 0; t < S; ++t) {
              int g = s + t;
              if (t >= nsl || g < glo || g > ghi) continue;
              uint64_t db = make_desc(sB + t * B_TILE, TN * 16, 128);
              tc_mma_i8(tbase + (uint32_t)((g - glo) * TN), da, db, idesc, (written >> g) & 1u);
              written |= 1u << g;
            }
          }
          tc_commit(&empty[st]);
        }
        tc_commit(tfull);
      }
    }
  } else {
    // epilogue: warp q = warp % 4 owns TMEM lanes 32q .. 32q+31 (rows of the tile)
    const int q = warp & 3;
    const int row = rb * TM + q * 32 + lane;
    const double srow = sa[row];
    double* crow = C + (long long)row * ldc + (long long)cb * TN;
#pragma unroll
    for (int r = 0; r < Cfg::NR; ++r) {
      const int glo = Cfg::glo(r), ghi = Cfg::ghi(r);
      mbar_wait(tfull, r & 1, 3);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = (r == 0) ? 0.0 : crow[c0 + j];   // partial Horner sum of the low groups
#pragma unroll
        for (int g = S - 1; g >= 0; --g) {        // smallest weight first: acc = acc / 128 + G_g
          if (g < glo || g > ghi) continue;
          uint32_t v[16];
          tc_ld16(tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)((g - glo) * TN + c0), v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fma(acc[j], 0.0078125, (double)(int)v[j]);
        }
        if (r == Cfg::NR - 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) crow[c0 + j] = acc[j] * srow * sb[cb * TN + c0 + j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) crow[c0 + j] = acc[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tempty)) : "memory");
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"((uint32_t)Cfg::TCOLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Issue-rate probe: one CTA per SM issues `count` kind::i8 MMAs of shape 128 x N x 32 from fixed shared-memory tiles
// into rotating accumulators; cycles per instruction -> the INT8 rate this instruction shape can reach.
template <int N>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int count, int same_a, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tbase = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    uint32_t sA = smem_u32(smem + 1024), sB = sA + 8 * A_TILE;
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      uint64_t da = make_desc(sA + (same_a ? 0 : (i & 7) * A_TILE), TM * 16, 128);
      uint64_t db = make_desc(sB + (i & 3) * (N * KSTEP), N * 16, 128);
      tc_mma_i8(tbase + (uint32_t)((i % (512 / N)) * N), da, db, idesc, 1u);
    }
    tc_commit(bar);
    mbar_wait(bar, 0, 9);
    long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory"); }
}

template <int N>
void run_rate(int sms, double ghz) {
  long long* d; CK(cudaMalloc(&d, sms * sizeof(long long)));
  int smem = 1024 + 8 * A_TILE + 4 * 256 * KSTEP;
  CK(cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int same = 0; same < 2; ++same) {
    const int count = 8192;
    mma_rate_kernel<N><<<sms, 64, smem>>>(count, same, d); CK(cudaDeviceSynchronize());
    mma_rate_kernel<N><<<sms, 64, smem>>>(count, same, d); CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    double mx = 0; for (auto v : h) mx = fmax(mx, (double)v);
    double cyc = mx / count;
    printf("  \"rate_M128_N%d_%s\": {\"cycles_per_mma\": %.1f, \"int8_tops_all_sms_at_%.3f_GHz\": %.0f},\n", N, same ? "sameA" : "rotA", cyc, ghz,
           2.0 * TM * N * KSTEP / cyc * ghz * sms * 1e-3);
  }
  CK(cudaFree(d));
}


// 2-D tensor map over a slice array seen as rows of 128-byte core matrices; box = one slice tile (32 rows = 4096 bytes).
// No swizzle / interleave: the box lands in shared memory exactly as it lies in global memory (the UMMA no-swizzle layout).
static CUtensorMap make_map(const void* base, size_t bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "cuTensorMapEncodeTiled not found\n"); exit(2); }
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {128, (cuuint64_t)(bytes / 128)};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {128, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(2); }
  return m;
}

// ------------------------------------------------------------------------------------------------------------------
struct Result { double ms_slice, ms_gemm, max_rel_norm, max_rel_comp; };

template <int S>
Result run_ozaki(const double* dA, const double* dB, double* dC, int M, int N, int K, int reps,
                 const std::vector<double>& hA, const std::vector<double>& hB, const std::vector<int>& si, const std::vector<int>& sj,
                 const std::vector<long double>& ref, const std::vector<long double>& refabs, double cmax) {
  using Cfg = GemmCfg<S>;
  int nk = K / KSTEP;
  int8_t *dAs, *dBs; double *dsa, *dsb;
  CK(cudaMalloc(&dAs, (size_t)M * K * S)); CK(cudaMalloc(&dBs, (size_t)N * K * S));
  CK(cudaMalloc(&dsa, M * sizeof(double))); CK(cudaMalloc(&dsb, N * sizeof(double)));
  CK(cudaFuncSetAttribute(ozaki_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  CUtensorMap mapA = make_map(dAs, (size_t)M * K * S), mapB = make_map(dBs, (size_t)N * K * S);
  cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  float ts = 0, tg = 0;
  for (int it = 0; it < reps + 1; ++it) {
    CK(cudaEventRecord(e0));
    slice_kernel<S, TM><<<(M + 7) / 8, 256>>>(dA, M, K, K, dAs, dsa);
    slice_kernel<S, TN><<<(N + 7) / 8, 256>>>(dB, N, K, K, dBs, dsb);
    CK(cudaEventRecord(e1));
    ozaki_gemm_kernel<S><<<dim3(N / TN, M / TM), 192, Cfg::SMEM>>>(mapA, mapB, dsa, dsb, dC, N, nk);
    CK(cudaEventRecord(e2));
    CK(cudaEventSynchronize(e2));
    CK(cudaGetLastError());
    if (it > 0) { float a, b; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2)); ts += a; tg += b; }
  }
  Result r; r.ms_slice = ts / reps; r.ms_gemm = tg / reps; r.max_rel_norm = 0; r.max_rel_comp = 0;
  for (size_t s = 0; s < si.size(); ++s) {
    double c; CK(cudaMemcpy(&c, dC + (size_t)si[s] * N + sj[s], sizeof(double), cudaMemcpyDeviceToHost));
    long double d = fabsl((long double)c - ref[s]);
    r.max_rel_norm = fmax(r.max_rel_norm, (double)(d / cmax));
    r.max_rel_comp = fmax(r.max_rel_comp, (double)(d / refabs[s]));
  }
  CK(cudaFree(dAs)); CK(cudaFree(dBs)); CK(cudaFree(dsa)); CK(cudaFree(dsb));
  return r;
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 4096;
  int K = argc > 2 ? atoi(argv[2]) : 4096;
  const char* ffile = argc > 3 ? argv[3] : nullptr;
  int M = n, N = n;
  if (M % TM || N % TN || K % KSTEP) { fprintf(stderr, "n must be a multiple of 128, K of 32\n"); return 1; }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  if (prop.major != 10) { fprintf(stderr, "needs sm_100 (found %d.%d)\n", prop.major, prop.minor); return 1; }
  printf("{\"device\": \"%s\", \"sms\": %d, \"M\": %d, \"N\": %d, \"K\": %d,\n", prop.name, prop.multiProcessorCount, M, N, K);
  if (getenv("OZAKI_RATE")) {
    double ghz = prop.clockRate * 1e-6;
    printf(" \"issue_rate\": {\n");
    run_rate<64>(prop.multiProcessorCount, ghz); run_rate<128>(prop.multiProcessorCount, ghz); run_rate<256>(prop.multiProcessorCount, ghz);
    printf("  \"end\": 0 },\n");
  }

  for (int test = 0; test < (ffile ? 3 : 2); ++test) {
    std::vector<double> hA((size_t)M * K), hB((size_t)N * K);
    const char* name;
    std::mt19937_64 rng(1234 + test);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    if (test == 0) { name = "uniform(-1,1)"; for (auto& v : hA) v = U(rng); for (auto& v : hB) v = U(rng); }
    else if (test == 1) {
      name = "wide dynamic range within rows: u * 10^(-6 u')";
      for (auto& v : hA) v = U(rng) * pow(10.0, -6.0 * fabs(U(rng)));
      for (auto& v : hB) v = U(rng) * pow(10.0, -6.0 * fabs(U(rng)));
    } else {
      // rows of a Cholesky factor written by tools/ozaki_gen.py: (M + N) x K doubles, row-major
      name = "Cholesky factor panel of an ArdSE Gram matrix (trailing-update operands)";
      FILE* f = fopen(ffile, "rb");
      if (!f || fread(hA.data(), 8, hA.size(), f) != hA.size() || fread(hB.data(), 8, hB.size(), f) != hB.size()) { fprintf(stderr, "cannot read %s\n", ffile); return 1; }
      fclose(f);
    }
    double *dA, *dB, *dC, *dCref;
    CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8));
    CK(cudaMalloc(&dC, (size_t)M * N * 8)); CK(cudaMalloc(&dCref, (size_t)M * N * 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));

    // cuBLAS DGEMM: C(row-major M x N) = A Bᵀ  <=>  column-major C' (N x M) = B'ᵀ-op ... : C' = op(B) op(A) with B' = K x N
    cublasHandle_t hb; cublasCreate(&hb);
    const double one = 1.0, zero = 0.0;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float tblas = 0; int reps = 3;
    for (int it = 0; it < reps + 1; ++it) {
      CK(cudaEventRecord(e0));
      cublasDgemm(hb, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, dB, K, dA, K, &zero, dCref, N);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      if (it > 0) { float a; CK(cudaEventElapsedTime(&a, e0, e1)); tblas += a; }
    }
    tblas /= reps;

    // long-double reference on a sample of entries
    const int NS = 512;
    std::vector<int> si(NS), sj(NS); std::vector<long double> ref(NS), refabs(NS);
    std::uniform_int_distribution<int> Ui(0, M - 1), Uj(0, N - 1);
    for (int s = 0; s < NS; ++s) {
      si[s] = Ui(rng); sj[s] = Uj(rng);
      long double acc = 0, aab = 0;
      const double* a = &hA[(size_t)si[s] * K]; const double* b = &hB[(size_t)sj[s] * K];
      for (int k = 0; k < K; ++k) { long double p = (long double)a[k] * (long double)b[k]; acc += p; aab += fabsl(p); }
      ref[s] = acc; refabs[s] = aab > 0 ? aab : 1;
    }
    double cmax = 0; for (int s = 0; s < NS; ++s) cmax = fmax(cmax, (double)fabsl(ref[s]));
    double blas_norm = 0, blas_comp = 0;
    for (int s = 0; s < NS; ++s) {
      double c; CK(cudaMemcpy(&c, dCref + (size_t)si[s] * N + sj[s], 8, cudaMemcpyDeviceToHost));
      long double d = fabsl((long double)c - ref[s]);
      blas_norm = fmax(blas_norm, (double)(d / cmax)); blas_comp = fmax(blas_comp, (double)(d / refabs[s]));
    }
    double flops = 2.0 * M * N * K;
    printf(" \"%s\": {\n  \"cublas_dgemm\": {\"ms\": %.3f, \"tflops\": %.2f, \"err_vs_max\": %.3e, \"err_vs_sum_abs\": %.3e},\n", name, tblas, flops / tblas * 1e-9, blas_norm, blas_comp);
    auto report = [&](int S, const Result& r) {
      printf("  \"ozaki_int8_S%d\": {\"slice_ms\": %.3f, \"gemm_ms\": %.3f, \"fp64_equiv_tflops_gemm\": %.2f, \"fp64_equiv_tflops_total\": %.2f, \"int8_tops\": %.1f, \"err_vs_max\": %.3e, \"err_vs_sum_abs\": %.3e},\n",
             S, r.ms_slice, r.ms_gemm, flops / r.ms_gemm * 1e-9, flops / (r.ms_slice + r.ms_gemm) * 1e-9, flops * (S * (S + 1) / 2) / r.ms_gemm * 1e-9, r.max_rel_norm, r.max_rel_comp);
      fflush(stdout);
    };
    report(5, run_ozaki<5>(dA, dB, dC, M, N, K, reps, hA, hB, si, sj, ref, refabs, cmax));
    report(6, run_ozaki<6>(dA, dB, dC, M, N, K, reps, hA, hB, si, sj, ref, refabs, cmax));
    report(7, run_ozaki<7>(dA, dB, dC, M, N, K, reps, hA, hB, si, sj, ref, refabs, cmax));
    report(8, run_ozaki<8>(dA, dB, dC, M, N, K, reps, hA, hB, si, sj, ref, refabs, cmax));
    printf("  \"end\": 0 },\n");
    cublasDestroy(hb);
    CK(cudaFree(dA)); CK(cudaFree(dB)); CK(cudaFree(dC)); CK(cudaFree(dCref));
  }
  printf(" \"done\": 1}\n");
  return 0;
}
