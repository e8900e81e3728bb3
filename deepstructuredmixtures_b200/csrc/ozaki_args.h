// Argument blocks of the INT8-tensor-core path of the triangular inverse (ozaki.cuh, api_ozaki.cu).
//
// X = L^-1 of one expert is split recursively:  X = [X11 0; X21 X22],  X21 = -X22 (L21 X11).  The diagonal parts stay
// on the tile pipeline (trtri3.cuh, FP64 DMMA); the two products per split are GEMMs without dependencies between
// their tiles and run on the INT8 tcgen05 tensor cores through an error-free Ozaki split of the FP64 operands
// (tools/ozaki_proto.cu is the stand-alone prototype with the accuracy / rate measurements).
#pragma once
#include <cstdint>
#include "common.cuh"

namespace dsm {

constexpr int OZ_KSTEP = 32;                 // K of one tcgen05.mma kind::i8
constexpr int OZ_TILE_B = BLK * OZ_KSTEP;    // bytes of one slice tile: 128 rows x 32 k, UMMA canonical K-major, no swizzle:
                                             //   [2 chunks of 16 k][16 groups of 8 rows][8 rows][16 bytes]
constexpr int OZ_KSTEPS_PER_BLK = BLK / OZ_KSTEP;

// Slice one 128 x 128 block of an operand.  `src` is a block in the factor-tile layout: (r, c) at
// (c >> 4) * TILE_D + (c & 15) * LDS + r  (8 consecutive factor tiles, or a W / W^T block, or a scratch block).
struct OzJob {
  const double* src;
  int transposed;        // operand(i, k) = src(k, i) instead of src(i, k)
  int vr, vk;            // valid operand rows / k entries; everything else is sliced as zero (and never read)
  int scale;             // index of operand row 0 of this block in the row-scale arrays
  int64_t dst;           // slice-pool tile index of (row block, first k-step of this block), slice 0
};

// One 128 x 128 output block:  out = sign * A[k0 .. k1) B[k0 .. k1)^T
struct OzTile {
  int a_tile, b_tile;    // slice-pool tile index of k-step 0, slice 0 of the A / B row block
  int k0, k1;            // k-steps
  int sa, sb;            // row-scale indices of the A rows (output rows) / B rows (output columns)
  int vr, vc;            // valid rows / columns of the output block
  double* out;           // factor-tile layout block
  double sign;
  int accum;             // 1: out += sign * product (right-looking update of the factorisation), 0: out = sign * product
  int pad_;
};

struct OzPart { int slot, I, J; };      // inverse tile (I, J) written by a GEMM: its fused partials are computed afterwards

struct OzPartArgs {
  const LeafMeta* meta;
  const double* F; const double* z;
  const int64_t* flag_off;
  double* apart; double* tpart;
  const OzPart* parts;
};

// host side plan of one batch (api_ozaki.cu)
struct OzLevel {
  OzJob* d_jobs1 = nullptr; int n_jobs1 = 0; OzTile* d_tiles1 = nullptr; int n_tiles1 = 0;     // stage 1: T^T = X11^T L21^T
  OzJob* d_jobs2 = nullptr; int n_jobs2 = 0; OzTile* d_tiles2 = nullptr; int n_tiles2 = 0;     // stage 2: X21^T = -T^T X22^T
  int n_scale = 0;
  double ksteps1 = 0, ksteps2 = 0;                   // sum over the block products of their k-steps (for the op counts)
};
struct OzPlan {
  bool active = false;
  void* d_blob = nullptr;                          // ONE device allocation that holds every list below
  int4* d_tasks = nullptr; int n_tasks = 0;        // tile-pipeline tasks restricted to the diagonal ranges
  OzLevel levels[4]; int n_levels = 0;             // deepest level first
  OzPart* d_parts = nullptr; int n_parts = 0;
  double gemm_flops = 0.0;
  double tile_flops = 0.0;                         // factorisation + inverse flops left on the FP64 tile pipelines (2/3 r^3 per diagonal range)
  // right-looking split of the factorisation at the root split: columns < mid (launch A), A22 -= L21 L21^T on the INT8 tensor
  // cores, columns >= mid with the contraction starting at mid (launch B).  The L21 slices are reused by the inverse.
  bool potrf = false;
  int4* d_potrfA = nullptr; int n_potrfA = 0; int4* d_potrfB = nullptr; int n_potrfB = 0;
  int* d_kskip = nullptr;
  OzJob* d_jobsL = nullptr; int n_jobsL = 0;         // L21 blocks
  OzTile* d_syrk = nullptr; int n_syrk = 0;
  int l21_scale0 = 0, l21_nscale = 0;
  // gradient evaluations: the tile-pipeline part of the inverse rides in the two factorisation launches (fused2.cuh):
  // tiles of the first diagonal range (and of unsplit experts) behind launch A, tiles of the second range behind launch B
  int4* d_invA = nullptr; int n_invA = 0; int4* d_invB = nullptr; int n_invB = 0;
  // ... and with X11 at hand after launch A, the panel below it is a product as well:  L21 = A21 X11^T  (instead of panel
  // tasks with a triangular solve): launch A then holds only the tiles of A11 (and the unsplit experts)
  bool trsm = false;
  int4* d_potrfA11 = nullptr; int n_potrfA11 = 0;
  OzJob* d_jobsT = nullptr; int n_jobsT = 0;         // A21 blocks and X11 blocks
  OzTile* d_tilesT = nullptr; int n_tilesT = 0;
  int nscaleT = 0;
  double ksteps_syrk = 0, ksteps_T = 0;
  // LAUUM contraction F^-1_IJ = sum_{K >= I} X_KI^T X_KJ of the split experts as block products on X^T (slices in the level region of
  // the pool, tiles into the scratch); lauum3_kernel then only runs its dK-trace epilogue on them
  bool lauum = false;
  OzJob* d_jobsX = nullptr; int n_jobsX = 0; OzTile* d_tilesW = nullptr; int n_tilesW = 0; int nscaleX = 0; double ksteps_W = 0;
  int64_t* d_pre_base = nullptr;
};

// launchers (k_ozaki.cu).  `map` is the CUtensorMap of the slice pool (128 opaque bytes, built by oz_make_map).
cudaError_t oz_init_kernels();
int oz_make_map(void* map128, const void* pool, size_t bytes, int box_rows);          // 0 on success
int oz_round_slices(int S, int r);                                        // slices round r of the block-product kernel loads per operand
void launch_oz_slice(int S, const OzJob* jobs, int njobs, int pass, unsigned long long* rowmax, double* scale, int8_t* pool, cudaStream_t st);
void launch_oz_gemm(int S, const void* maps256, const OzTile* tiles, int ntiles, const double* scale, int nctas, cudaStream_t st, long long* trace = nullptr);
void launch_oz_parts(const OzPartArgs& a, int nparts, cudaStream_t st);
void launch_oz_setflags(const OzPart* parts, int n, const int64_t* flag_off, int* flags, cudaStream_t st);

}  // namespace dsm
