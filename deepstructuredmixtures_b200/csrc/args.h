// Kernel argument blocks and host-side launchers (one translation unit per kernel family so the library
// builds in parallel; api.cu sees only this header).
#pragma once
#include "common.cuh"

namespace dsm {

constexpr int GT = 64;       // gram tile
constexpr int GDC = 16;      // dimensions staged per pass



struct SolveArgs {
  const LeafMeta* meta;
  const double* F;
  const double* WT;
  const double* z;
  double* alpha;
  int* flags; const int64_t* flag_off;    // flag J of an expert: alpha_J stored
  const int2* tasks; int ntasks;          // (slot, J), descending J inside an expert
  int* counter; int* gerr;
  const int4* share;                      // per slot sharing plan (see ShareKind) or null: aliased experts are skipped
};

// Sharing plan of fit! (fit.jl:71-122), one int4 per slot: x = kind, y = source slot (the "main" expert), z = number of
// leading 128-row blocks whose factor tiles are copied from the source, w = unused.
//   SHARE_NONE   the expert is factored on its own (update_cholesky!)
//   SHARE_ALIAS  fitcontained!(Val(true), Val(true)) fit.jl:132-143: identical observations -> every result of the source is
//                reused (no Gram, no factorisation, no inverse; rows / alpha / predictions read the source's slot)
//   SHARE_PREFIX fitcontained! :145-292 (row deletion / chol_continue!): the experts share their first 128*z observations,
//                the factor tiles of those block rows are copied and the factorisation continues behind them
enum ShareKind : int { SHARE_NONE = 0, SHARE_ALIAS = 1, SHARE_PREFIX = 2 };

struct GramArgs {
  const LeafMeta* meta;
  const double* xg;
  const double* prm;
  double* F;
  const int64_t* tile_off;   // [nleaves+1] prefix sum of lower-triangular GT tiles per leaf
  int nleaves;
  int D;
  const int4* share;         // per slot sharing plan or null: aliased experts and copied block rows are not built
};

struct GramRectArgs {
  int ktype, D;
  const double* prm;
  const double* xa; int64_t sa; int na;
  const double* xb; int64_t sb; int nb;
  double* out; int64_t ldo;
};

struct GatherArgs {
  const LeafMeta* meta;
  const double* x; int64_t N; int D;
  const int64_t* obs;        // concatenated 1-based rows (local leaves)
  const int64_t* obs_off;    // [nleaves+1]
  double* xg;
};

struct LauumArgs {
  const LeafMeta* meta;
  const double* F;
  const double* WT;
  const double* xg;
  const double* alpha;
  const double* prm;
  const int4* tasks;        // (leaf slot, I, J, index of this task within the leaf)
  int ntasks;
  int* counter;
  double* gpart;            // [gpart_off[slot] + task_in_leaf * nl + h]
  const int64_t* gpart_off;
  int D;
  int* gerr;
  const int* mask;          // per slot: 0 = skip this expert, or null
  // F^-1 tiles already computed (INT8 block products, api_ozaki.cu): tile of task w of slot s = pre + (pre_base[s] + w) * WBLK_D, holding
  // -F^-1_IJ in the factor-tile layout (rows of block I); pre_base[s] < 0: this expert's tiles are contracted here.  null: all are
  const double* pre; const int64_t* pre_base;
};

struct RowsArgs {
  const LeafMeta* meta;
  const LeafScal* scal_in;
  LeafScal* scal;
  const double* prm;
  const double* trpart; const int64_t* trpart_off;
  const double* gpart; const int64_t* gpart_off;   // may be null when no LAUUM pass ran
  double* rows; int row_width;
  int as_written; int with_grad; int lauum_ran;
  const double* ldpart; const double* zzpart;   // engine v2: per block-column partials of logdet and z'z (else null)
  const double* alpha;                          // alpha'alpha is reduced here
  const int* mask;                              // per slot: 0 = gradient entries are written as 0 (expert skipped)
  const int4* share;                            // per slot sharing plan or null: aliased experts get their row copied afterwards
};

struct PredLeaf {     // per leaf with routed points
  int32_t slot;       // leaf slot (LeafMeta index)
  int32_t T;          // routed points
  int32_t Tp;         // padded to BLK
  int32_t pad_;
  int64_t xtoff;      // xt: D columns of length Tp
  int64_t vtoff;      // V^T scratch: Tp/128 row blocks of nkc tiles (factor tile layout)
  int64_t ooff;       // output offset (mu / var), length Tp
};

struct PredArgs {
  const LeafMeta* meta;
  const PredLeaf* pl;
  const int2* tasks;       // (pred leaf index, Q)
  int ntasks;
  int* counter;
  const double* F;
  const double* W;
  const double* xg;
  const double* alpha;
  const double* prm;
  const double* leaf_mean;  // per global leaf
  const double* xt;
  double* VT;
  double* mu;
  double* var;
  int D;
  int* gerr;
  // wave mode (few test points): one task per (pred leaf, Q, I), cross-CTA order by flags, per-task partials
  int wave;
  const int4* wtasks;      // (pred leaf index, Q, I, base): flag / partial slot of (pl, Q, K) = base + K
  int* flags;
  double* part;            // [slot][2][BLK]: mean partial, sum-of-squares partial
  const int4* wcols;       // (pred leaf index, Q, base, nb) per (pl, Q), for the final reduction
  int nwcols;
  // device-routed test points (large batches): the points of (expert, position) are gathered through pidx from the caller's
  // T x D matrix instead of a packed copy; position = PredLeaf.ooff + q, pidx < 0 on the padding
  const int* pidx;         // or null: xt holds the packed points
  const double* xtest; int64_t T_all;
  // V^T scratch per CTA (a task's V^T row block is written and read back by the CTA that owns the task only)
  int vt_per_cta; int64_t vt_stride;
  long long* trace;        // optional [ntasks][8] cycle counts per phase (DSMGP_PTRACE_FILE), else null
};

// ---- launchers (defined next to their kernels) ------------------------------------------------
void launch_solve(const SolveArgs& a, int nctas, cudaStream_t st);
int solve_max_ctas(int sms);
void launch_lauum3(const LauumArgs& a, int nctas, cudaStream_t st);
void launch_rows(const RowsArgs& a, int nleaves, cudaStream_t st);
void launch_predict3(const PredArgs& a, int nctas, cudaStream_t st);
void launch_predict_reduce(const PredArgs& a, cudaStream_t st);
void launch_gram_fit(const GramArgs& a, int64_t ntiles, cudaStream_t st);
void launch_gram_rect(const GramRectArgs& a, cudaStream_t st);
void launch_gather(const GatherArgs& a, int maxnp, int nleaves, cudaStream_t st);
struct Potrf2Args;
struct Trtri3Args;
cudaError_t init_v2_kernels();
void launch_potrf2(const Potrf2Args& a, int nctas, cudaStream_t st);
void launch_trtri3(const Trtri3Args& a, int nctas, const int2* cols, int ncols, cudaStream_t st);
void launch_trtri3_only(const Trtri3Args& a, int nctas, cudaStream_t st);                      // without the block-column reduction
void launch_alpha_reduce(const Trtri3Args& a, const int2* cols, int ncols, cudaStream_t st);
void launch_eval2(const Potrf2Args& pa, const Trtri3Args& ta, int nctas, const int2* cols, int ncols, cudaStream_t st);
void launch_eval2_only(const Potrf2Args& pa, const Trtri3Args& ta, int nctas, cudaStream_t st);      // without the block-column reduction
void launch_untile(const double* Ft, int nkc, int n, double* out, cudaStream_t st);
void launch_ov_count(const int64_t* obs, int64_t total, int* cntp, cudaStream_t st);
void launch_ov_fill(const int64_t* obs, const int64_t* leaf_ptr, int L, const int64_t* poff, int* fill, int* plist, cudaStream_t st);
void launch_ov_pairs(const int64_t* poff, const int* plist, int64_t N, int64_t L, int* inter, cudaStream_t st);
void launch_ov_finish(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD, const int* node_type,
                      int64_t L, double* D, cudaStream_t st);
void launch_ov_csr(const int* inter, const int64_t* leaf_ptr, const int* kid, const int* anc, int AD, const int* node_type, int64_t L,
                   int* row_cnt, const int64_t* row_ptr, int32_t* col, double* val, cudaStream_t st);
// jobs_dev: device array of { double* L; int n; const int64_t* rows; int nrows; double* v } (k_misc.cu DelJob), one CTA per matrix,
// at most DEL_QMAX = 64 deleted rows per job and launch
void launch_delete_rows(const void* jobs_dev, int njobs, cudaStream_t st);
struct RouteArgs;
struct MixArgs;
void launch_route(const RouteArgs& a, bool fill, cudaStream_t st);
void launch_mix(const MixArgs& a, cudaStream_t st);
void launch_share_copy(const LeafMeta* meta, const int4* share, const int* slots, int nslots, int max_jb, double* F, cudaStream_t st);
void launch_rows_alias(const LeafMeta* meta, const int4* share, int nslots, double* rows, int row_width, LeafScal* scal, cudaStream_t st);

}  // namespace dsm
