// Argument blocks of the engine-v2 kernels (host-visible).
#pragma once
#include "common.cuh"
namespace dsm {
struct Potrf2Args {
  const LeafMeta* meta;
  double* F; double* W; double* WT;
  const double* y; double* z;
  LeafScal* scal;
  double* trpart; const int64_t* trpart_off;      // [2*nb] per leaf: diag-block partials of tr(F^-1) in the first half
  double* ldpart; double* zzpart;                  // [nb] per leaf at trpart_off/2: logdet and z'z partials
  int* flags; const int64_t* flag_off;             // per leaf nb(nb+1)/2 tile flags
  const int4* tasks; int ntasks;                   // (slot, I, J, unused)
  int* counter; int* gerr;
  int jstart;                                      // chol_continue: block columns < jstart hold a valid factor
  const int4* share;                               // per slot sharing plan or null: aliased experts are skipped, SHARE_PREFIX
                                                   // experts continue behind their copied block rows (per-slot jstart = share.z)
  long long* trace;                                // optional [ntasks][8] clock stamps (DSMGP_TRACE_FILE), else null
  const int* kskip;                                // per slot block index ks or null: for block columns >= ks the k-blocks < ks have
                                                   // already been subtracted from the tiles (right-looking SYRK on the INT8 tensor
                                                   // cores, api_ozaki.cu), the contraction starts at ks
};

// tile-pipelined inverse (trtri3): tasks (slot, I, J, unused), I > J, ordered by anti-diagonal
struct Trtri3Args {
  const LeafMeta* meta;
  double* F; const double* W; const double* WT;
  const double* z; double* alpha;
  double* trpart; const int64_t* trpart_off;    // second half [nb + J], written by alpha_reduce_kernel
  int* flags; const int64_t* flag_off;          // per leaf nb(nb+1)/2 tile flags / partial slots
  double* apart; double* tpart;                 // per-tile partials: [tile][BLK], [tile]
  const int4* tasks; int ntasks;
  int* counter; int* gerr;
  const int* mask;                              // per slot: 0 = skip this expert (no gradient wanted), or null
};

}  // namespace dsm
